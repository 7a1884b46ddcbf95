"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatement of the reference's batch assembly for adaptive
bottom-up features -- dataset.py:288-355 (`split_entries` + `trim_collate`) -- in plain NumPy loops, small cases only.

Reference semantics restated:
  * dataset.py:302-304   per entry, rows [pos_boxes[img][0], pos_boxes[img][1]) of image_features / spatial_features / image_bb
  * dataset.py:314-318   target = zeros(num_ans); np.put_along_axis(target, labels, scores, 0)   (a later duplicate label wins)
  * dataset.py:334,340,346  keras pad_sequences(..., padding='post', maxlen=<longest in the batch>, dtype=float32):
                         zero rows appended after the real ones, every sample padded to the batch maximum
Pinned by tests/golden/refexec_collate.npz: the reference's own dataset.py (tensorize / split_entries / trim_collate) executed on
an in-memory store over oracle/tf_shim (oracle/make_golden_ref_collate.py).
Keras is not installed here; `pad_sequences` with padding='post' and maxlen equal to the longest sequence neither truncates
nor reorders, it only appends zeros (keras/utils/sequence_utils.py), which is what `_pad_post` does."""
import numpy as np


def _pad_post(seqs, dtype=np.float32):
    maxlen = max(s.shape[0] for s in seqs)
    out = np.zeros((len(seqs), maxlen) + seqs[0].shape[1:], dtype=dtype)
    for i, s in enumerate(seqs):
        for r in range(s.shape[0]):
            out[i, r] = s[r]
    return out


def collate(image_features, spatial_features, image_bb, pos_boxes, image_ids, labels, scores, num_ans):
    feats, nbbs, bbs, targets = [], [], [], []
    for k, img in enumerate(image_ids):
        lo, hi = int(pos_boxes[img][0]), int(pos_boxes[img][1])
        feats.append(image_features[lo:hi, :])
        nbbs.append(spatial_features[lo:hi, :])
        bbs.append(image_bb[lo:hi, :])
        t = np.zeros(num_ans)
        if labels[k] is not None:
            np.put_along_axis(t, np.asarray(labels[k]), np.asarray(scores[k], dtype=np.float64), 0)
        targets.append(t)
    return _pad_post(feats), _pad_post(nbbs), _pad_post(bbs), np.array(targets, dtype=np.float32)
