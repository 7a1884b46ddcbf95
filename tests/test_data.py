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


def test_ragged_collate_packs_the_same_rows():
    feats, nbb, bb, pos = _store(seed=2)
    st = AdaptiveFeatureStore(feats, nbb, bb, pos)
    ids = [5, 3, 3, 0, 8]
    rg = st.collate_ragged(ids)
    f0, _, b0, _ = oc.collate(feats, nbb, bb, pos, ids, [None] * 5, [None] * 5, 3)
    off = rg["offsets"].numpy()
    assert rg["offsets"].dtype == torch.int32 and off[0] == 0 and rg["max_rois"] == 100
    assert rg["features"].shape[0] == off[-1] == sum(pos[i, 1] - pos[i, 0] for i in ids)
    for k, i in enumerate(ids):                      # un-padding the reference's batch gives the packed rows back
        c = pos[i, 1] - pos[i, 0]
        np.testing.assert_array_equal(rg["features"][off[k]:off[k + 1]].numpy(), f0[k, :c])
        np.testing.assert_array_equal(rg["boxes"][off[k]:off[k + 1]].numpy(), b0[k, :c])
    out = st.ragged_buffers(5, 500, pin=False)
    rg2 = st.collate_ragged(ids[:2], out=out)        # reusable buffers
    assert rg2["features"].shape[0] == rg2["offsets"][-1] and rg2["features"].data_ptr() == out["features"].data_ptr()
    with pytest.raises(ValueError, match="too small"):
        st.collate_ragged([3] * 6, out=out)


def test_pad_on_device_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tf_vqa_regat_b200._lib import RegatError
    from tf_vqa_regat_b200.data import pad_on_device
    feats, nbb, bb, pos = _store()
    with pytest.raises(RegatError):
        pad_on_device(AdaptiveFeatureStore(feats, nbb, bb, pos).collate_ragged([0, 1]), "cuda:0")


@pytest.mark.gpu
def test_pad_on_device_is_bit_exact_with_host_padding():
    from tf_vqa_regat_b200.data import pad_on_device
    feats, nbb, bb, pos = _store(seed=3, images=12, V=2048)
    st = AdaptiveFeatureStore(feats, nbb, bb, pos)
    ids = [5, 3, 3, 0, 8, 11, 1]
    want = st.collate(ids)
    got = pad_on_device(st.collate_ragged(ids), "cuda:0")
    assert torch.equal(got["features"].cpu(), want["features"]) and torch.equal(got["boxes"].cpu(), want["boxes"])
    assert torch.equal(got["n_obj"].cpu(), want["n_obj"])
    # shorter batch, wider padding, into reused (dirty) output buffers: nothing stale survives
    out = {"features": torch.full((7, 100, 2048), 7.0, device="cuda:0"), "boxes": torch.full((7, 100, 4), 7.0, device="cuda:0")}
    ids2 = [0, 1]
    got2 = pad_on_device(st.collate_ragged(ids2), "cuda:0", pad_to=100, out=out)
    want2 = st.collate(ids2, pad_to=100)
    assert torch.equal(got2["features"].cpu(), want2["features"]) and torch.equal(got2["boxes"].cpu(), want2["boxes"])
    # invalid offsets (a sample longer than N) are flagged, never read
    rg = st.collate_ragged(ids)
    rg["max_rois"] = 100
    rg["offsets"] = rg["offsets"].clone()
    rg["offsets"][1] = -5
    with pytest.raises(ValueError, match="invalid"):
        pad_on_device(rg, "cuda:0")


@pytest.mark.gpu
def test_engine_on_device_padded_batch_matches_host_padded_batch():
    """The hot path sees the same bits either way: logits are identical for a ragged-shipped and a host-padded batch."""
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.config import HotPathConfig
    from tf_vqa_regat_b200.data import pad_on_device
    from tf_vqa_regat_b200.engine import HotPathEngine
    cfg = HotPathConfig(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)
    rng = np.random.default_rng(5)
    counts = np.array([36, 12, 20, 29]); ends = np.cumsum(counts); T = int(ends[-1])
    feats = np.maximum(rng.standard_normal((T, 192)), 0).astype(np.float32)
    bb = np.sort(rng.uniform(0, 400, (T, 4)), axis=1).astype(np.float32)
    st = AdaptiveFeatureStore(feats, np.zeros((T, 6), np.float32), bb, np.stack([ends - counts, ends], 1))
    ids = [1, 0, 3, 2]
    host = st.collate(ids)
    devp = pad_on_device(st.collate_ragged(ids), "cuda:0")
    eng = HotPathEngine(cfg, 4, 36, dtype="fp32")
    eng.load_params(syn.make_params(cfg, seed=7, trained_like=True))
    q = torch.tensor(rng.standard_normal((2, 4, 96)).astype(np.float32)).cuda()
    a = eng.forward(host["features"].cuda(), host["boxes"].cuda(), q[0], q[1])
    b = eng.forward(devp["features"], devp["boxes"], q[0], q[1])
    assert torch.equal(a, b)


def test_collate_matches_the_executed_reference(golden_dir):
    """tests/golden/refexec_collate.npz: the reference's own dataset.py (tensorize / split_entries / trim_collate) executed on
    an in-memory store (oracle/make_golden_ref_collate.py).  Host collate, ragged collate and the oracle restatement all
    reproduce it bit for bit."""
    import os
    g = np.load(os.path.join(golden_dir, "refexec_collate.npz"))
    st = AdaptiveFeatureStore(g["image_features"], g["spatial_features"], g["image_bb"], g["pos_boxes"])
    ids = g["ids"].tolist()
    parse = lambda s, t: [t(x) for x in str(s).split(",")] if str(s) else []
    labels = [parse(s, int) for s in g["labels"]]
    scores = [parse(s, float) for s in g["scores"]]
    got = st.collate(ids)
    assert got["features"].dtype == torch.float32
    np.testing.assert_array_equal(got["features"].numpy(), g["features"])
    np.testing.assert_array_equal(got["normalized_bb"].numpy(), g["normalized_bb"])
    np.testing.assert_array_equal(got["boxes"].numpy(), g["boxes"])
    t = targets_from_answers(labels, scores, int(g["num_ans"])).numpy()
    np.testing.assert_array_equal(t, g["targets"])
    assert t[0, 4] == np.float32(0.9) and t[5, 1] == np.float32(0.9) and not t[1].any()      # last duplicate wins; no answers -> zeros
    f0, n0, b0, t0 = oc.collate(g["image_features"], g["spatial_features"], g["image_bb"], g["pos_boxes"], ids,
                                [l if l else None for l in labels], scores, int(g["num_ans"]))
    np.testing.assert_array_equal(f0, g["features"]); np.testing.assert_array_equal(n0, g["normalized_bb"])
    np.testing.assert_array_equal(b0, g["boxes"]); np.testing.assert_array_equal(t0, g["targets"])
    rg = st.collate_ragged(ids)
    off = rg["offsets"].numpy()
    for k in range(len(ids)):
        c = off[k + 1] - off[k]
        np.testing.assert_array_equal(rg["features"][off[k]:off[k + 1]].numpy(), g["features"][k, :c])
        assert not g["features"][k, c:].any()


def test_tokenize_questions_matches_the_executed_reference(golden_dir):
    """tests/golden/refexec_tokens.json: Dictionary.tokenize + VQAFeatureDataset.tokenize of the reference's dataset.py, executed."""
    import json
    import os
    from tf_vqa_regat_b200.data import tokenize_questions
    g = json.load(open(os.path.join(golden_dir, "refexec_tokens.json")))
    t = tokenize_questions(g["questions"], g["word2idx"])
    assert t.dtype == torch.int32 and tuple(t.shape) == (len(g["questions"]), 14)
    assert t.tolist() == g["q_token"]
    assert g["padding_idx"] == g["ntoken"] and (t[6] == g["ntoken"]).all()                # the empty question is all padding
    assert t[0, 4].item() == g["word2idx"]["'s"]                                           # "man's" -> "man", "'s"
    unk = g["ntoken"] - 1
    assert unk in t[7].tolist() and len([x for x in t[7].tolist() if x != g["ntoken"]]) == 14   # unknown -> last word; cut at 14
    buf = torch.zeros(16, 14, dtype=torch.int32)
    assert tokenize_questions(g["questions"][:2], g["word2idx"], out=buf).data_ptr() == buf.data_ptr()
