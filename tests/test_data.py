"""Batch assembly (tf_vqa_regat_b200/data.py) against the restatement of dataset.py:288-355 (oracle/dataset_collate.py)."""
import numpy as np
import pytest
import torch

from oracle import dataset_collate as oc
from tf_vqa_regat_b200.data import AdaptiveFeatureStore, targets_from_answers


def _store(seed=0, images=9, V=24):
    rng = np.random.default_rng(seed)
    counts = rng.integers(10, 101, size=images)
    counts[3] = 100                                  # the adaptive maximum
    ends = np.cumsum(counts)
    pos = np.stack([ends - counts, ends], axis=1)
    T = int(ends[-1])
    return (rng.standard_normal((T, V)).astype(np.float32), rng.random((T, 6)).astype(np.float32),
            (rng.random((T, 4)) * 600).astype(np.float32), pos)


def test_collate_is_bit_exact_with_the_reference_restatement():
    feats, nbb, bb, pos = _store()
    st = AdaptiveFeatureStore(feats, nbb, bb, pos)
    ids = [5, 3, 3, 0, 8]                            # repeated image, arbitrary order
    labels = [[4, 7, 4], None, [0], [], [2, 9]]       # duplicate label: the later score wins (np.put_along_axis)
    scores = [[0.3, 1.0, 0.9], None, [0.6], [], [0.3, 0.3]]
    f0, n0, b0, t0 = oc.collate(feats, nbb, bb, pos, ids, [l if l else None for l in labels], scores, num_ans=11)
    got = st.collate(ids)
    assert got["features"].shape == (5, 100, 24) and got["features"].dtype == torch.float32
    np.testing.assert_array_equal(got["features"].numpy(), f0)
    np.testing.assert_array_equal(got["normalized_bb"].numpy(), n0)
    np.testing.assert_array_equal(got["boxes"].numpy(), b0)
    np.testing.assert_array_equal(got["n_obj"].numpy(), [pos[i, 1] - pos[i, 0] for i in ids])
    np.testing.assert_array_equal(targets_from_answers(labels, scores, 11).numpy(), t0)
    assert float(targets_from_answers(labels, scores, 11)[0, 4]) == pytest.approx(0.9)


def test_reused_buffers_leave_no_stale_rows_and_pad_to():
    feats, nbb, bb, pos = _store(seed=1)
    st = AdaptiveFeatureStore(feats, nbb, bb, pos)
    out = st.buffers(4, 100, pin=False)
    st.collate([3, 3, 3, 3], out=out)                # fills all 100 rows of every sample
    got = st.collate([0, 1], out=out, pad_to=100)    # shorter samples into the same buffers
    f0, _, b0, _ = oc.collate(feats, nbb, bb, pos, [0, 1], [None, None], [None, None], 3)
    n = f0.shape[1]
    np.testing.assert_array_equal(got["features"][:, :n].numpy(), f0)
    assert not got["features"][:, n:].any() and not got["boxes"][:, n:].any()       # zero post-padding, nothing stale
    np.testing.assert_array_equal(got["boxes"][:, :n].numpy(), b0)
    with pytest.raises(ValueError, match="pad_to"):
        st.collate([3], pad_to=50)
    with pytest.raises(ValueError, match="too small"):
        st.collate([0, 1, 2, 4, 5], out=out)


def test_store_validates_shapes():
    feats, nbb, bb, pos = _store()
    with pytest.raises(ValueError):
        AdaptiveFeatureStore(feats, nbb[:-1], bb, pos)
    with pytest.raises(ValueError):
        AdaptiveFeatureStore(feats, nbb, bb, pos.reshape(-1))
