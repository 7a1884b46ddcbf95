"""Question front-end on the GPU (tf_vqa_regat_b200/question.py + csrc/question.cu): each kernel against its NumPy
statement (tests/_host_emulation.py), the front-end against the oracle, and the WHOLE model -- tokens in, logits out,
gradients into the embedding tables, two optimizer steps -- against the vectors produced by executing the reference's own
language_model.py / rel_graph_net.py / train.py (tests/golden/refexec_question_*.npz)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import language_model as olm
from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig

from _host_emulation import HostOps

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _both(fn_name, arrays, scalars_fn, outs):
    """Run one entry point on the device and its emulation on the host with the same inputs; return {name: (dev, host)}."""
    from tf_vqa_regat_b200 import _lib
    L, H = _lib.lib(), HostOps()
    dev = {k: (torch.tensor(v).cuda() if v is not None else None) for k, v in arrays.items()}
    host = {k: (np.ascontiguousarray(v).copy() if v is not None else None) for k, v in arrays.items()}
    dp = {k: (t.data_ptr() if t is not None else None) for k, t in dev.items()}
    hp = {k: (a.ctypes.data if a is not None else None) for k, a in host.items()}
    _lib.check(getattr(L, fn_name)(*scalars_fn(dp), torch.cuda.current_stream().cuda_stream))
    assert getattr(H, fn_name)(*scalars_fn(hp), None) == 0
    torch.cuda.synchronize()
    return {k: (dev[k].cpu().numpy(), host[k]) for k in outs}


def test_embed_kernels():
    rng = np.random.default_rng(0)
    n_token, E, BT = 30, 12, 5 * 14
    tok = rng.integers(0, n_token + 1, BT).astype(np.int32)
    tok[:3] = n_token
    emb, emb2 = rng.standard_normal((n_token + 1, E)).astype(np.float32), rng.standard_normal((n_token + 1, E)).astype(np.float32)
    r = _both("regat_q_embed_fwd", dict(tok=tok, emb=emb, emb2=emb2, out=np.zeros((BT, 2 * E), np.float32)),
              lambda p: (p["tok"], BT, n_token, E, p["emb"], p["emb2"], p["out"]), ["out"])
    assert np.array_equal(*r["out"]) and not r["out"][0][:3].any()
    dX = rng.standard_normal((BT, 2 * E)).astype(np.float32)
    r = _both("regat_q_embed_bwd", dict(tok=tok, dX=dX, d1=np.zeros_like(emb), d2=np.zeros_like(emb2)),
              lambda p: (p["tok"], BT, n_token, E, 2 * E, p["dX"], p["d1"], p["d2"]), ["d1", "d2"])
    for k in ("d1", "d2"):
        np.testing.assert_allclose(*r[k], rtol=1e-5, atol=1e-5)
        assert not r[k][0][n_token].any()


def test_gru_gate_kernels():
    rng = np.random.default_rng(1)
    B, H, T = 5, 24, 3
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    z = lambda *s: np.zeros(s, np.float32)
    arrs = dict(xi=f(B, T * 3 * H), hi=f(B, 3 * H), hp=f(B, T * H), h=z(B, T * H), zz=z(B, H), rr=z(B, H), cc=z(B, H), hpc=z(B, H))
    r = _both("regat_q_gru_gates_fwd", arrs,
              lambda p: (B, H, p["xi"] + 4 * 3 * H, T * 3 * H, p["hi"], p["hp"], T * H, p["h"] + 4 * H, T * H, p["zz"], p["rr"], p["cc"], p["hpc"]),
              ["h", "zz", "rr", "cc", "hpc"])
    for k, (d, h) in r.items():
        np.testing.assert_allclose(d, h, rtol=2e-6, atol=2e-6, err_msg=k)
    zz, rr, cc, hpc = (r[k][1] for k in ("zz", "rr", "cc", "hpc"))
    arrs = dict(dseq=f(B, T * H), drec=f(B, H), zz=zz, rr=rr, cc=cc, hpc=hpc, hi=arrs["hi"], dxi=z(B, T * 3 * H), dhi=z(B, 3 * H), dhp=z(B, H))
    r = _both("regat_q_gru_gates_bwd", arrs,
              lambda p: (B, H, p["dseq"] + 4 * H, T * H, p["drec"], p["zz"], p["rr"], p["cc"], p["hpc"], p["hi"], p["dxi"] + 4 * 3 * H, T * 3 * H,
                         p["dhi"], p["dhp"]), ["dxi", "dhi", "dhp"])
    for k, (d, h) in r.items():
        np.testing.assert_allclose(d, h, rtol=1e-5, atol=2e-6, err_msg=k)


def test_attention_kernels():
    rng = np.random.default_rng(2)
    B, T, H = 7, 14, 24
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    r = _both("regat_q_batch_softmax_fwd", dict(lg=3 * f(B, T), P=np.zeros((T, B), np.float32)), lambda p: (p["lg"], B, T, p["P"]), ["P"])
    np.testing.assert_allclose(*r["P"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r["P"][0].sum(1), 1.0, rtol=1e-5)             # normalised over the batch, per position
    P = r["P"][1]
    r = _both("regat_q_batch_softmax_bwd", dict(P=P, dP=f(T, B), dl=np.zeros((B, T), np.float32)), lambda p: (p["P"], p["dP"], B, T, p["dl"]), ["dl"])
    np.testing.assert_allclose(*r["dl"], rtol=1e-5, atol=1e-6)
    seq = f(B, T, H)
    r = _both("regat_q_pool_fwd", dict(P=P, seq=seq, q=np.zeros((B, H), np.float32)), lambda p: (p["P"], p["seq"], B, T, H, p["q"]), ["q"])
    np.testing.assert_allclose(*r["q"], rtol=1e-5, atol=1e-6)
    r = _both("regat_q_pool_bwd", dict(P=P, seq=seq, ga=f(B, H), gl=f(B, H), ds=np.zeros((B, T, H), np.float32), dW=np.zeros((B, T), np.float32)),
              lambda p: (p["P"], p["seq"], p["ga"], p["gl"], B, T, H, p["ds"], p["dW"]), ["ds", "dW"])
    np.testing.assert_allclose(*r["ds"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(*r["dW"], rtol=1e-5, atol=1e-5)
    x = f(1000)
    r = _both("regat_q_tanh_fwd", dict(x=x), lambda p: (p["x"], 1000), ["x"])
    np.testing.assert_allclose(*r["x"], rtol=1e-6, atol=1e-6)
    r = _both("regat_q_tanh_bwd", dict(dy=f(1000), y=r["x"][1]), lambda p: (p["dy"], p["y"], 1000), ["dy"])
    np.testing.assert_allclose(*r["dy"], rtol=1e-5, atol=1e-6)


def test_reduction_weight_norm_and_optimizer_kernels():
    rng = np.random.default_rng(3)
    n = 5000
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    a, b = f(n), f(n)
    r = _both("regat_q_dot", dict(a=a, b=b, o=np.zeros(1, np.float32)), lambda p: (p["a"], p["b"], n, p["o"]), ["o"])
    np.testing.assert_allclose(*r["o"], rtol=1e-4)
    g, vv, Gv = np.array([1.7], np.float32), np.array([float(b @ b)], np.float32), np.array([float(a @ b)], np.float32)
    r = _both("regat_q_wn_alpha", dict(g=g, vv=vv, al=np.zeros(1, np.float32)), lambda p: (p["g"], p["vv"], p["al"]), ["al"])
    np.testing.assert_allclose(*r["al"], rtol=1e-6)
    r = _both("regat_q_wn_bwd", dict(G=a, v=b, g=g, vv=vv, Gv=Gv, dv=np.zeros(n, np.float32), dg=np.zeros(1, np.float32)),
              lambda p: (p["G"], p["v"], p["g"], p["vv"], p["Gv"], n, p["dv"], p["dg"]), ["dv", "dg"])
    np.testing.assert_allclose(*r["dv"], rtol=1e-5, atol=1e-6); np.testing.assert_allclose(*r["dg"], rtol=1e-5)
    w, m, u = f(n), 0.1 * f(n), np.abs(f(n))
    ss = np.array([float(a @ a)], np.float32)
    r = _both("regat_q_clip_adamax", dict(w=w, g=a, m=m, u=u, ss=ss),
              lambda p: (p["w"], p["g"], p["m"], p["u"], n, p["ss"], 0.25, 1e-3, 3, 0.9, 0.999, 1e-8), ["w", "m", "u"])
    for k in ("w", "m", "u"):
        np.testing.assert_allclose(*r[k], rtol=1e-5, atol=1e-7, err_msg=k)


def test_embedding_sparse_clip_and_adamax_kernels():
    """The IndexedSlices path of train.py:112-113 for the embedding tables: occurrence-norm (regat_q_embed_sumsq) and the sparse
    Adamax (regat_q_embed_clip_adamax) against the host emulation AND the oracle statement pinned to the executed reference."""
    rng = np.random.default_rng(8)
    n_token, E, BT, W = 7, 12, 40, 24                       # 7 words over 40 positions: every word repeats; table in columns [12, 24)
    tok = rng.integers(0, n_token + 1, BT).astype(np.int32)
    dX = rng.standard_normal((BT, W)).astype(np.float32)
    r = _both("regat_q_embed_sumsq", dict(t=tok, d=dX, o=np.zeros(1, np.float32)), lambda p: (p["t"], BT, n_token, E, W, 12, p["d"], p["o"]), ["o"])
    np.testing.assert_allclose(*r["o"], rtol=1e-5)
    want_ss = float((dX[tok != n_token][:, 12:].astype(np.float64) ** 2).sum())
    np.testing.assert_allclose(r["o"][0][0], want_ss, rtol=1e-5)
    n = (n_token + 1) * E
    table, m, u = rng.standard_normal(n).astype(np.float32), 0.1 * rng.standard_normal(n).astype(np.float32), np.abs(rng.standard_normal(n)).astype(np.float32) * 0.05
    dense = np.zeros((n_token + 1, E), np.float32)
    np.add.at(dense, tok[tok != n_token], dX[tok != n_token][:, 12:])
    ss = np.array([want_ss], np.float32)
    arrays = dict(t=tok, d=dX, w=table, g=dense.ravel(), m=m, u=u, ui=np.zeros(n, np.float32), ss=ss)
    r = _both("regat_q_embed_clip_adamax", arrays,
              lambda p: (p["t"], BT, n_token, E, W, 12, p["d"], p["w"], p["g"], p["m"], p["u"], p["ui"], p["ss"], 0.25, 1e-3, 3, 0.9, 0.999, 1e-8),
              ["w", "m", "u", "ui"])
    for k in ("w", "m", "u"):
        np.testing.assert_allclose(*r[k], rtol=1e-5, atol=1e-7, err_msg=k)
    assert not r["ui"][0].any()
    from oracle import language_model as olm
    w2, m2, u2 = olm.sparse_clip_adamax(table.astype(np.float64).reshape(-1, E), tok, dX[:, 12:], m.astype(np.float64).reshape(-1, E),
                                         u.astype(np.float64).reshape(-1, E), 3, 1e-3, 0.25, 0.9, 0.999, 1e-8, n_token)
    np.testing.assert_allclose(r["w"][0].reshape(-1, E), w2, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r["u"][0].reshape(-1, E), u2, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("op,B,emb2_tr", [("c", 6, True), ("", 3, False)])
def test_front_end_matches_oracle(op, B, emb2_tr):
    from tf_vqa_regat_b200.question import QuestionFrontEnd
    n_token, E, H, T = 60, 12, 96, 14
    fe = QuestionFrontEnd(n_token, E, H, op=op, seq_len=T, max_batch=8, emb2_trainable=emb2_tr)
    front = olm.make_params(n_token, E, H, op, seed=11)
    fe.load_named(front)
    tok = olm.make_tokens(B, n_token, T, seed=21)
    q_att, q_last = fe.forward(torch.tensor(tok, dtype=torch.int32).cuda())
    rng = np.random.default_rng(5)
    dqa, dql = rng.standard_normal((B, H)).astype(np.float32), rng.standard_normal((B, H)).astype(np.float32)
    p = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in front.items()}
    q = olm.forward(p, tok, n_token, op)
    ((q["q_att"] * torch.tensor(dqa, dtype=torch.float64)).sum() + (q["q_last"] * torch.tensor(dql, dtype=torch.float64)).sum()).backward()
    assert _rel(q_att.cpu().numpy(), q["q_att"].detach().numpy()) < 1e-4
    assert _rel(q_last.cpu().numpy(), q["q_last"].detach().numpy()) < 1e-4
    fe.backward(torch.tensor(dqa).cuda(), torch.tensor(dql).cuda())
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in fe.named(fe.grads).items()}
    for name, g in got.items():
        if name == "w_emb.emb_/emb_" and not emb2_tr:
            assert not g.any()
            continue
        if name == "q_att.linear2/bias":
            continue                                                     # softmax-shift direction: rounding noise on both sides
        assert _rel(g, p[name].grad.numpy()) < 5e-4, name


def test_whole_model_tokens_to_logits_matches_reference_execution():
    """tokens -> front-end -> hot path -> loss -> gradients back into the embedding tables -> clip + Adamax on every variable,
    two steps, against the executed reference (refexec_question_small)."""
    from tf_vqa_regat_b200.engine import HotPathEngine
    from tf_vqa_regat_b200.question import QuestionFrontEnd
    g = np.load(os.path.join(HERE, "golden", "refexec_question_small.npz"))
    cfg = HotPathConfig(**ast.literal_eval(str(g["cfg"])))
    B, N, steps, lr = int(g["B"]), int(g["N"]), int(g["steps"]), float(g["lr"])
    n_token, E, op = int(g["n_token"]), int(g["emb_dim"]), str(g["op"])
    batches = [syn.make_inputs(cfg, B, N, seed=1000 + s, adaptive=bool(g["adaptive"])) for s in range(steps + 1)]
    np.testing.assert_allclose([float(np.sum(b["features"], dtype=np.float64)) for b in batches], g["input_check"], rtol=1e-12)
    eng = HotPathEngine(cfg, B, N, dtype="fp32")
    eng.load_params(syn.make_params(cfg, seed=7, trained_like=True))
    fe = QuestionFrontEnd(n_token, E, cfg.q_dim, op=op, seq_len=14, max_batch=B, emb2_trainable=bool(g["emb2_trainable"]),
                          grad_clip=cfg.grad_clip, beta1=cfg.beta1, beta2=cfg.beta2, eps=cfg.eps)
    fe.load_named(olm.make_params(n_token, E, cfg.q_dim, op, seed=11))
    for s in range(steps):
        d = {k: torch.tensor(v).cuda() for k, v in batches[s].items() if k != "n_obj"}
        tok = torch.tensor(g["tokens"][s], dtype=torch.int32).cuda()
        q_att, q_last = fe.forward(tok)
        out = eng.fwd_bwd(d["features"], d["boxes"], q_att, q_last, d["target"], want_dq=True, want_logits=True)
        fe.backward(out["dq_att"], out["dq_last"])
        torch.cuda.synchronize()
        if s == 0:
            for k, want in (("q_att", q_att), ("q_last", q_last)):
                assert _rel(want.cpu().numpy(), g[k]) < 1e-4, k
            assert _rel(out["logits"].cpu().numpy(), g["logits"]) < 1e-4
            assert abs(float(out["loss"]) - float(g["loss"])) < 1e-4 * float(g["loss"])
            got = {k: v.cpu().numpy() for k, v in fe.named(fe.grads).items()}
            for name, gr in got.items():
                if "grad/" + name in g.files and name != "q_att.linear2/bias":
                    assert _rel(gr, g["grad/" + name]) < 1e-3, name
        assert _rel(out["logits"].cpu().numpy(), g["train.logits"][s]) < (1e-4 if s == 0 else 5e-3)
        eng.update(lr, s + 1)
        fe.update(lr, s + 1)
    got = {k: v.cpu().numpy() for k, v in fe.named().items()}
    for name, w in got.items():
        if name == "q_att.linear2/bias":
            continue
        assert np.abs(w - g["param/" + name]).max() < 0.05 * steps * lr + 1e-6, name
