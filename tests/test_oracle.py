"""The oracle against (a) golden vectors produced by the reference's own position_emb.py,
(b) its own committed fp64 outputs, (c) a second independent transcription (torch),
(d) the fused re-association the kernels use, (e) finite differences."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from oracle import position_emb as ope
from oracle import regat_fused as ofu
from oracle import regat_numpy as onp
from oracle import regat_torch as ot
from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig, num_trainable, param_layout

TINY = dict(v_dim=96, q_dim=48, rel_dim=64, num_heads=4, nongt_dim=5, num_answers=37)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stage1_*.npz"))))
def test_stage1_bit_exact_vs_reference_golden(path):
    g = np.load(path)
    nongt = int(g["nongt"])
    pm = ope.extract_position_matrix(g["bb"], nongt_dim=nongt)
    emb, a, b = ope.prepare_graph_variables("implicit", g["bb"], None, None, g["bb"].shape[1], nongt, 64, 11, 15)
    assert a is None and b is None
    assert pm.dtype == g["pos_mat"].dtype and np.array_equal(pm, g["pos_mat"])
    assert emb.shape == g["pos_emb"].shape and np.array_equal(emb.astype(np.float32), g["pos_emb"])
    assert np.isfinite(emb).all()                      # padded (0,0,0,0) boxes stay finite


def test_stage1_feature_index_and_scramble():
    rng = np.random.default_rng(3)
    bb = np.sort(rng.uniform(0, 400, (1, 36, 4)).astype(np.float32), axis=-1)[:, :, [0, 1, 2, 3]]
    bb = np.stack([bb[..., 0], bb[..., 1], bb[..., 2], bb[..., 3]], -1)
    pm = ope.extract_position_matrix(bb, 20)
    emb = ope.extract_position_embedding(pm, 64)
    div = ope.wave_divisors(64)
    # feature c*16+k = sin(100*P_c/div_k), c*16+8+k = cos
    for c, k in [(0, 0), (1, 3), (2, 7), (3, 5)]:
        arg = (np.float32(100.0) * pm[0, 4, 9, c]) / div[k]
        assert emb[0, 4, 9, c * 16 + k] == np.sin(arg) and emb[0, 4, 9, c * 16 + 8 + k] == np.cos(arg)
    ii, jj = ope.scrambled_pair_index(36, 20)
    assert (ii[3, 7], jj[3, 7]) == (1, 31) and (ii[35, 19], jj[35, 19]) == (19, 35)
    ii, jj = ope.scrambled_pair_index(12, 20)
    assert np.array_equal(ii, np.arange(12)[:, None].repeat(12, 1)) and np.array_equal(jj, np.arange(12)[None].repeat(12, 0))


def _case(kw, B, N, adaptive, tl, dtype=np.float64):
    cfg = HotPathConfig(**kw)
    inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=adaptive)
    named = syn.unflatten(cfg, syn.make_params(cfg, seed=7, trained_like=tl).astype(dtype))
    c = lambda a: a.astype(dtype)
    return cfg, inp, named, (c(inp["features"]), inp["boxes"], c(inp["q_att"]), c(inp["q_last"]), c(inp["target"]))


@pytest.mark.parametrize("name", ["tiny_n9_m5", "tiny_n4_m5", "full_b2_n36_m20"])
def test_hotpath_golden_regression(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"hotpath_{name}.npz"))
    kw = ast.literal_eval(str(g["cfg"]))
    cfg, inp, named, args = _case(kw, int(g["B"]), int(g["N"]), bool(g["adaptive"]), bool(g["trained_like"]))
    out = onp.forward(named, cfg, *args)
    np.testing.assert_allclose(out["logits"], g["logits"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["loss"], g["loss"], rtol=1e-12)
    np.testing.assert_array_equal(out["mask"], g["mask"])
    np.testing.assert_allclose(out["v1"].sum(-1), g["v1_sum"], rtol=1e-10)


@pytest.mark.parametrize("N,adaptive,tl", [(9, True, True), (4, False, False), (5, False, True)])
def test_numpy_vs_torch_vs_fused(N, adaptive, tl):
    cfg, inp, named, args = _case(TINY, 3, N, adaptive, tl)
    a = onp.forward(named, cfg, *args)
    b = ot.forward(ot.to_torch_params(named, requires_grad=False), cfg, *args)
    c = ofu.forward(named, cfg, *args)
    for k in ("logits", "v1", "joint", "att_weights"):
        np.testing.assert_allclose(a[k], b[k].numpy(), rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(a[k], c[k], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(a["loss"], c["loss"], rtol=1e-12)
    np.testing.assert_array_equal(a["mask"], c["mask"])
    for d in range(cfg.dir_num):                        # intermediates the kernels save
        np.testing.assert_allclose(a["att"][d]["prob"].transpose(0, 2, 1, 3), c["dirs"][d]["P"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(a["att"][d]["z"].transpose(0, 2, 1, 3), c["dirs"][d]["z"], rtol=1e-9, atol=1e-12)


def test_mask_semantics_padded_rows():
    # At init (bias 0) padded rows have v0 == 0 -> masked; with trained-like biases relu(b0) != 0 -> unmasked (A.2-Q8)
    cfg, inp, named, args = _case(TINY, 4, 9, True, False)
    out = onp.forward(named, cfg, *args)
    expect = (np.arange(9)[None, :] < inp["n_obj"][:, None]).astype(np.float64)
    np.testing.assert_array_equal(out["mask"], expect)
    cfg, inp, named, args = _case(TINY, 4, 9, True, True)
    assert onp.forward(named, cfg, *args)["mask"].all()


def test_full_dims_fp32_close_to_fp64():
    cfg, inp, named, args = _case({}, 2, 36, False, True)
    named32 = {k: v.astype(np.float32) for k, v in named.items()}
    a = onp.forward(named, cfg, *args)
    b = onp.forward(named32, cfg, *[x.astype(np.float32) for x in args])
    assert b["logits"].dtype == np.float32
    err = np.abs(a["logits"] - b["logits"]).max() / np.abs(a["logits"]).max()
    assert err < 1e-4, err


def test_autograd_matches_finite_differences():
    cfg, inp, named, args = _case(TINY, 2, 6, False, True)
    loss, grads, dq_att, dq_last, _ = ot.loss_and_grads(named, cfg, inp)
    rng = np.random.default_rng(0)
    for name in ["v_relation.v2out/v", "v_relation.implicit_relation.neighbor_net.1.pair_pos_fc/v",
                 "v_relation.implicit_relation.neighbor_net.0.linear_out_/v", "joint_emb.linear/g",
                 "v_relation.implicit_relation.neighbor_net.0.key/bias", "classifier.layers.3/g"]:
        w = named[name]
        idx = tuple(rng.integers(0, s) for s in w.shape)
        h = 1e-6 * max(1.0, abs(float(w[idx])))
        orig = float(w[idx])
        vals = []
        for sgn in (+1, -1):
            w[idx] = orig + sgn * h
            vals.append(float(onp.forward(named, cfg, *args)["loss"]))
        w[idx] = orig
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - grads[name][idx]) <= 1e-5 * max(1.0, abs(fd)), (name, fd, grads[name][idx])
    # label-FC gradient is softmax-shift noise (A.2-Q9)
    assert np.abs(grads["v_relation.implicit_relation.bias/v"]).max() < 1e-12


def test_clip_and_adamax_restatement():
    g = np.array([3.0, 4.0])
    np.testing.assert_allclose(ot.clip_by_norm(g, 0.25), g * 0.05)
    np.testing.assert_allclose(ot.clip_by_norm(g * 0.01, 0.25), g * 0.01)
    w, m, u = ot.adamax_step(np.ones(2), np.array([0.5, -2.0]), np.zeros(2), np.zeros(2), 1, 1e-3)
    np.testing.assert_allclose(m, [0.05, -0.2]); np.testing.assert_allclose(u, [0.5, 2.0])
    np.testing.assert_allclose(w, 1 - (1e-3 / 0.1) * m / (u + 1e-8))


def test_param_layout_order_and_count():
    cfg = HotPathConfig()
    entries, total = param_layout(cfg)
    assert num_trainable(cfg) == 18_980_717
    names = [e.name for e in entries]
    assert names[:3] == ["v_relation.v2out/v", "v_relation.v2out/g", "v_relation.v2out/bias"]
    assert "v_relation.implicit_relation.bias/bias" not in names      # label_bias False
    assert names[-3:] == ["classifier.layers.3/v", "classifier.layers.3/g", "classifier.layers.3/bias"]
    assert all(e.offset % 64 == 0 for e in entries) and total % 64 == 0
    i = names.index
    assert i("v_relation.implicit_relation.neighbor_net.0.pair_pos_fc/v") < i("v_relation.implicit_relation.neighbor_net.0.query/v") \
        < i("v_relation.implicit_relation.neighbor_net.0.key/v") < i("v_relation.implicit_relation.neighbor_net.0.linear_out_/v") \
        < i("v_relation.implicit_relation.neighbor_net.1.pair_pos_fc/v") < i("joint_emb.v2attention/v")
