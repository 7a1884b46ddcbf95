"""The host-side mirror of the reference's layer API (tf_vqa_regat_b200/model) against the oracle and the fused engine."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import regat_numpy as onp
from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig, param_layout

pytestmark = pytest.mark.gpu
SMALL = dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stage1_*.npz"))))
def test_prepare_graph_variables_vs_reference_golden(path):
    """Stage 1 on the GPU against the outputs of the reference's own position_emb.py."""
    from tf_vqa_regat_b200.model.position_emb import prepare_graph_variables
    g = np.load(path)
    nongt = int(g["nongt"])
    pos_emb, a, b = prepare_graph_variables("implicit", g["bb"], None, None, g["bb"].shape[1], nongt, 64, 11, 15)
    assert a is None and b is None and tuple(pos_emb.shape) == g["pos_emb"].shape
    got = pos_emb.cpu().numpy()
    # sin/cos arguments reach |x| ~ 690 where one fp32 ulp is 6e-5: logf/sincosf may differ from NumPy's by an ulp
    assert np.abs(got - g["pos_emb"]).max() < 3e-4
    assert (np.abs(got - g["pos_emb"]) < 1e-6).mean() > 0.9
    assert np.isfinite(got).all()


@pytest.mark.parametrize("N,nongt,lazy", [(36, 20, True), (36, 20, False), (12, 20, True), (36, 36, False)])
def test_layer_model_matches_oracle_and_engine(N, nongt, lazy):
    from tf_vqa_regat_b200.engine import HotPathEngine
    from tf_vqa_regat_b200.model import build_hot_path, prepare_graph_variables
    cfg = HotPathConfig(**dict(SMALL, nongt_dim=nongt))
    B = 3
    inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=True)
    flat = syn.make_params(cfg, seed=7, trained_like=True)
    model = build_hot_path(cfg)
    names = [n for n, _ in model.weights]
    assert len(names) == len(param_layout(cfg)[0])
    assert [n.split("/")[-1] for n in [e.name for e in param_layout(cfg)[0]]] == names      # v, g, bias per layer, same order
    model.load_flat(cfg, flat)
    np.testing.assert_array_equal(model.to_flat(cfg), flat)
    dev = {k: torch.tensor(v).cuda() for k, v in inp.items() if k != "n_obj"}
    pos_emb, _, _ = prepare_graph_variables("implicit", dev["boxes"], None, None, N, cfg.nongt_dim, cfg.pos_emb_dim, 11, 15, lazy=lazy)
    logits = model(dev["features"], dev["q_att"], dev["q_last"], pos_emb)
    named64 = syn.unflatten(cfg, flat.astype(np.float64))
    f64 = lambda a: a.astype(np.float64)
    ref = onp.forward(named64, cfg, f64(inp["features"]), inp["boxes"], f64(inp["q_att"]), f64(inp["q_last"]))
    assert _rel(logits.cpu().numpy(), ref["logits"]) < 1e-4
    eng = HotPathEngine(cfg, B, N, dtype="fp32", training=False)
    eng.load_params(flat)
    l2 = eng.forward(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"])
    assert _rel(logits.cpu().numpy(), l2.cpu().numpy()) < 2e-5


def test_single_attention_layer_and_butd_vs_oracle():
    from tf_vqa_regat_b200.model import BUTD, GraphSelfAttentionLayer, prepare_graph_variables
    from tf_vqa_regat_b200.model import _rt
    cfg = HotPathConfig(**dict(SMALL, rel_dim=512, num_heads=8))      # one direction: num_heads must be a multiple of 8
    B, N, D, H = 2, 36, cfg.rel_dim, cfg.num_heads
    rng = np.random.default_rng(5)
    roi = rng.standard_normal((B, N, D)).astype(np.float32)
    inp = syn.make_inputs(cfg, B, N, seed=3)
    layer = GraphSelfAttentionLayer(D, cfg.nongt_dim, pos_emb_dim=64, num_heads=H)
    pos_emb, _, _ = prepare_graph_variables("implicit", inp["boxes"], None, None, N, cfg.nongt_dim, 64, 11, 15)
    lab = torch.full((B, N, cfg.nongt_dim), 0.37, device="cuda")
    out = layer(torch.tensor(roi).cuda(), torch.ones(B, N, cfg.nongt_dim, device="cuda"), pos_emb, lab)
    names = ["pair_pos_fc", "query", "key", "linear_out_"]
    ws = layer.get_weights()
    assert len(ws) == 12 and ws[9].shape == (1, 1, D, D)                      # variable order of graph_att_layer.py:25-37
    p = {}
    for i, n in enumerate(names):
        p[f"L.{n}/v"], p[f"L.{n}/g"], p[f"L.{n}/bias"] = (w.astype(np.float64) for w in ws[3 * i:3 * i + 3])
    ref, _ = onp.graph_self_attention_layer(p, "L", roi.astype(np.float64), np.ones((B, N, cfg.nongt_dim)),
                                            pos_emb.cpu().numpy().astype(np.float64), np.full((B, N, cfg.nongt_dim), 0.37), cfg)
    assert _rel(out.cpu().numpy(), ref) < 1e-4
    # BUTD
    butd = BUTD(D, cfg.q_dim, cfg.q_dim)
    v = torch.tensor(roi).cuda(); q = torch.tensor(inp["q_last"]).cuda()
    joint, w = butd(v, q)
    ws = butd.get_weights()
    p = {}
    for i, n in enumerate(["v2attention", "q2attention", "linear", "visual_embed", "question_embed"]):
        p[f"joint_emb.{n}/v"], p[f"joint_emb.{n}/g"], p[f"joint_emb.{n}/bias"] = (x.astype(np.float64) for x in ws[3 * i:3 * i + 3])
    rj, rw = onp.butd(p, roi.astype(np.float64), inp["q_last"].astype(np.float64))
    assert tuple(w.shape) == (B, N, 1) and _rel(w.cpu().numpy(), rw) < 1e-4 and _rel(joint.cpu().numpy(), rj) < 1e-4
    assert _rel(butd.attention_weights(v, q).cpu().numpy(), rw) < 1e-4


def test_reference_error_behaviour():
    from tf_vqa_regat_b200._lib import RegatError
    from tf_vqa_regat_b200.model import GraphAttentionNetwork, WeightNorm
    from tf_vqa_regat_b200.model.relation_encoder import concat_visual_question
    net = GraphAttentionNetwork(2, 1, 352, 256, nongt_dim=20, num_heads=4, pos_emb_dim=64)
    x = torch.zeros(1, 4, 352, device="cuda")
    with pytest.raises(ValueError):                      # graph_att_net.py:42-46
        net(x, None, None)
    net2 = GraphAttentionNetwork(2, 1, 352, 256, nongt_dim=20, num_heads=4, pos_emb_dim=-1)
    with pytest.raises(ValueError):                      # graph_att_net.py:47-51
        net2(x, None, torch.zeros(1, 4, 4, 64, device="cuda"))
    with pytest.raises(ValueError):                      # weight_norm.py:12-13
        WeightNorm(object())
    with pytest.raises(RegatError):                      # no CPU path
        concat_visual_question(torch.zeros(1, 8), torch.zeros(1, 2, 8))
    # mask semantics, bit exact (relation_encoder.py:20-21)
    v = torch.zeros(2, 3, 8, device="cuda"); v[0, 0, 2] = 1.0; v[1, 2, 0] = -2.0; v[1, 1, 0] = 1.0; v[1, 1, 1] = -1.0
    q = torch.arange(16, dtype=torch.float32, device="cuda").view(2, 8) + 1
    out = concat_visual_question(q, v).cpu().numpy()
    expect_mask = np.array([[1, 0, 0], [0, 0, 1]], dtype=np.float32)          # row (1,1) sums to exactly 0 -> masked
    np.testing.assert_array_equal(out[..., 8:], expect_mask[..., None] * q.cpu().numpy()[:, None, :])
    np.testing.assert_array_equal(out[..., :8], v.cpu().numpy())


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_layer_model_trains_through_the_engine(dtype):
    """The training path behind the layer API (train.py:103-113 tapes model(...)): model.compile binds the layers' own variables
    as the engine's parameter buffer, model.train_step runs tape + clip + Adamax, and afterwards the layer-by-layer forward,
    the compiled forward and get_weights all show the updated weights."""
    from tf_vqa_regat_b200.engine import HotPathEngine
    from tf_vqa_regat_b200.model import build_hot_path, prepare_graph_variables
    cfg = HotPathConfig(**SMALL)
    B, N, lr = 4, 36, 2e-3
    inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=True)
    flat = syn.make_params(cfg, seed=7, trained_like=True)
    dev = {k: torch.tensor(v).cuda() for k, v in inp.items() if k != "n_obj"}
    model = build_hot_path(cfg)
    model.load_flat(cfg, flat)
    model.compile(cfg, B, N, dtype=dtype)
    np.testing.assert_array_equal(model.to_flat(cfg), flat)                 # rebinding kept the values
    geo, _, _ = prepare_graph_variables("implicit", dev["boxes"], None, None, N, cfg.nongt_dim, cfg.pos_emb_dim, 11, 15, lazy=True)
    ref = HotPathEngine(cfg, B, N, dtype=dtype)
    ref.load_params(flat)
    before = model(dev["features"], dev["q_att"], dev["q_last"], geo).clone()
    for s in range(3):
        l = model.train_step(dev["features"], dev["q_att"], dev["q_last"], geo, dev["target"], lr, s + 1)
        w = ref.train_step(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"], dev["target"], lr, s + 1)
        assert abs(float(l[0]) - float(w[0])) < (2e-5 if dtype == "fp32" else 2e-3) * abs(float(w[0]))
    # one set of weights: the layers see what the engine trained
    got = model.to_flat(cfg)
    dd = np.abs(got - ref.params.cpu().numpy())
    # two engines, same arithmetic, but atomics order noise decides the sign of a gradient that is ~0 and Adamax moves such an element by
    # +-lr per step: single elements may differ by up to 2 lr per step, the mean stays far below
    assert dd.max() <= 2 * lr * 3 + 1e-6 and dd.mean() < 0.02 * lr
    assert np.abs(got - flat).max() > 0.5 * lr                                # and they really moved
    layerwise = model(dev["features"], dev["q_att"], dev["q_last"], geo)     # fp32 layer-by-layer kernels on the trained weights
    compiled = model.predict(dev["features"], dev["q_att"], dev["q_last"], geo)
    assert _rel(layerwise.cpu().numpy(), compiled.cpu().numpy()) < (2e-5 if dtype == "fp32" else 1e-2)
    assert _rel(layerwise.cpu().numpy(), before.cpu().numpy()) > 1e-3
    # set_weights through the layer API reaches the engine's derived state (alpha, bf16 kernels)
    model.load_flat(cfg, flat)
    again = model.predict(dev["features"], dev["q_att"], dev["q_last"], geo)
    fresh = HotPathEngine(cfg, B, N, dtype=dtype); fresh.load_params(flat)
    want = fresh.forward(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"])
    assert torch.equal(again, want)
    out = model.train_step(dev["features"], dev["q_att"], dev["q_last"], geo, dev["target"], lr, 1, want_dq=True)
    assert out[1][0].shape == (B, cfg.q_dim) and out[1][1].shape == (B, cfg.q_dim)
    with pytest.raises(TypeError):
        model.predict(dev["features"], dev["q_att"], dev["q_last"], torch.zeros(B, 20, N, 64, device="cuda"))
