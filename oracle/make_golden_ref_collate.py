"""Golden vectors for the batch assembly of adaptive bottom-up features, produced by executing the reference's
own dataset.py: VQAFeatureDataset.tensorize / split_entries / trim_collate (dataset.py:270-355) on an in-memory feature store.
Build container only:

    python -m oracle.make_golden_ref_collate        # rewrites tests/golden/refexec_collate.npz

The dataset object is created without running its constructor (which opens HDF5 / pickle files that do not exist here); the
attributes the collate path reads are set by hand: features / normalized_bb / bb / pos_boxes (dataset.py:207-230), entries,
num_ans_candidates.  Stand-ins: tensorflow (oracle/tf_shim; pad_sequences restated from Keras), h5py (import only).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    sys.path.insert(0, "/root/reference")
    import dataset as ref_dataset
    rng = np.random.default_rng(7)
    images, V, A = 9, 24, 11
    counts = rng.integers(10, 101, size=images)
    counts[3] = 100
    ends = np.cumsum(counts)
    pos_boxes = np.stack([ends - counts, ends], axis=1)
    T = int(ends[-1])
    feats = rng.standard_normal((T, V)).astype(np.float32)
    nbb = rng.random((T, 6)).astype(np.float32)
    bb = (rng.random((T, 4)) * 600).astype(np.float32)
    ids = [5, 3, 3, 0, 8, 1]
    labels = [[4, 7, 4], [], [0], [2, 9], [10], [1, 1, 1]]                 # duplicates: np.put_along_axis keeps the last score
    scores = [[0.3, 1.0, 0.9], [], [0.6], [0.3, 0.3], [1.0], [0.3, 0.6, 0.9]]
    tokens = rng.integers(0, 50, (len(ids), 14))

    ds = object.__new__(ref_dataset.VQAFeatureDataset)
    ds.features, ds.normalized_bb, ds.bb, ds.pos_boxes = feats, nbb, bb, pos_boxes
    ds.semantic_adj_matrix = ds.spatial_adj_matrix = None
    ds.num_ans_candidates = A
    ds.entries = [{"image": int(i), "q_token": list(map(int, t)), "answer": {"labels": list(l), "scores": list(s)}}
                  for i, t, l, s in zip(ids, tokens, labels, scores)]
    ds.tensorize()                                                        # dataset.py:270-286
    ds.batch_entries = [ds.entries]
    f, n, q, b, spa, sem, tgt = ds.split_entries(0)                       # dataset.py:288-355
    out = dict(image_features=feats, spatial_features=nbb, image_bb=bb, pos_boxes=pos_boxes, ids=np.array(ids), tokens=tokens,
               labels=np.array([",".join(map(str, l)) for l in labels]), scores=np.array([",".join(map(str, s)) for s in scores]),
               num_ans=A, features=f, normalized_bb=n, boxes=b, questions=q.numpy(), targets=tgt.numpy())
    assert f.dtype == np.float32 and f.shape == (len(ids), 100, V)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refexec_collate.npz"), **out)
    print("refexec collate:", f.shape, n.shape, b.shape, tgt.shape, "targets dtype", tgt.numpy().dtype)


if __name__ == "__main__":
    main()
