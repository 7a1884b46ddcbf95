"""Batch assembly for adaptive bottom-up features (reference dataset.py:206-230 layout, :288-355 collate).

The reference keeps the whole feature store in host memory -- `image_features [T, V]`, `spatial_features [T, 6]`,
`image_bb [T, 4]` with `pos_boxes[img] = (first row, one past last row)` -- and builds every batch with Python lists and
`pad_sequences` (zero post-padding to the longest sample, dataset.py:334-346).  Here the same result is written straight
into reusable (optionally pinned) buffers with one contiguous block copy per sample, ready for a single host->device copy;
the padded rows are exactly zero, which is what the hot path's padded-object handling relies on (relation_encoder.py:20-21,
SURVEY A.2-Q7/Q8)."""
from typing import Dict, Optional, Sequence

import numpy as np
import torch


class AdaptiveFeatureStore:
    def __init__(self, image_features: np.ndarray, spatial_features: np.ndarray, image_bb: np.ndarray, pos_boxes: np.ndarray):
        if not (image_features.shape[0] == spatial_features.shape[0] == image_bb.shape[0]):
            raise ValueError("image_features, spatial_features and image_bb must have the same number of rows")
        if pos_boxes.ndim != 2 or pos_boxes.shape[1] != 2:
            raise ValueError("pos_boxes must be [num_images, 2]")
        self.features = np.ascontiguousarray(image_features, dtype=np.float32)
        self.normalized_bb = np.ascontiguousarray(spatial_features, dtype=np.float32)
        self.bb = np.ascontiguousarray(image_bb, dtype=np.float32)
        self.pos_boxes = np.asarray(pos_boxes, dtype=np.int64)

    def counts(self, image_ids: Sequence[int]) -> np.ndarray:
        pb = self.pos_boxes[np.asarray(image_ids, dtype=np.int64)]
        return (pb[:, 1] - pb[:, 0]).astype(np.int64)

    def buffers(self, batch: int, max_rois: int, pin: bool = True) -> Dict[str, torch.Tensor]:
        """Reusable host buffers for `collate(out=...)`; pinned when CUDA is available so the H2D copy is asynchronous."""
        pin = bool(pin and torch.cuda.is_available())
        mk = lambda *s: torch.zeros(*s, dtype=torch.float32, pin_memory=pin)
        return {"features": mk(batch, max_rois, self.features.shape[1]), "normalized_bb": mk(batch, max_rois, self.normalized_bb.shape[1]),
                "boxes": mk(batch, max_rois, self.bb.shape[1]), "n_obj": torch.zeros(batch, dtype=torch.int64)}

    def collate(self, image_ids: Sequence[int], out: Optional[Dict[str, torch.Tensor]] = None, pad_to: Optional[int] = None):
        """-> dict(features [B,N,V], normalized_bb [B,N,6], boxes [B,N,4], n_obj [B]) with N = longest sample of the batch
        (dataset.py:334) or `pad_to`.  With `out` (from `buffers`) nothing is allocated and views of `out` are returned."""
        ids = np.asarray(image_ids, dtype=np.int64)
        n = self.counts(ids)
        B, N = len(ids), int(n.max()) if len(ids) else 0
        if pad_to is not None:
            if pad_to < N:
                raise ValueError(f"pad_to={pad_to} is shorter than the longest sample ({N} objects)")
            N = pad_to
        if out is None:
            out = self.buffers(B, N, pin=False)
        for k in ("features", "normalized_bb", "boxes"):
            if out[k].shape[0] < B or out[k].shape[1] < N:
                raise ValueError(f"buffer '{k}' {tuple(out[k].shape)} is too small for a batch of {B} x {N}")
        f, nb, bb = (out[k][:B, :N].numpy() for k in ("features", "normalized_bb", "boxes"))
        for i, img in enumerate(ids):
            lo, hi = self.pos_boxes[img]
            c = hi - lo
            f[i, :c] = self.features[lo:hi]; f[i, c:] = 0.0
            nb[i, :c] = self.normalized_bb[lo:hi]; nb[i, c:] = 0.0
            bb[i, :c] = self.bb[lo:hi]; bb[i, c:] = 0.0
        out["n_obj"][:B] = torch.from_numpy(n)
        return {"features": out["features"][:B, :N], "normalized_bb": out["normalized_bb"][:B, :N], "boxes": out["boxes"][:B, :N],
                "n_obj": out["n_obj"][:B]}


def targets_from_answers(labels, scores, num_answers: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dataset.py:314-318: soft-score targets [B, num_answers]; a later duplicate label overwrites an earlier one
    (np.put_along_axis); entries with labels None stay zero."""
    B = len(labels)
    t = out[:B] if out is not None else torch.zeros(B, num_answers, dtype=torch.float32)
    t.zero_()
    a = t.numpy()
    for i, (l, s) in enumerate(zip(labels, scores)):
        if l is not None and len(l):
            a[i, np.asarray(l, dtype=np.int64)] = np.asarray(s, dtype=np.float32)
    return t
