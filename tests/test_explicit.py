"""Explicit relation encoders.  CPU: the oracle restatement and the vectorised build_graph against vectors produced by
executing the reference's own files (oracle/make_golden_ref_explicit.py).  GPU: the layer mirror's ExplicitRelationEncoder against
the same vectors, and the masked attention kernels (forward, backward, label-bias reduction) against a torch fp32 reference."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from tf_vqa_regat_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "refexec_explicit_*.npz")))
IDS = [os.path.basename(f)[len("refexec_explicit_"):-4] for f in FILES]


def _case(path):
    g = np.load(path)
    cfg = ast.literal_eval(str(g["cfg"]))
    B, N, seed = int(g["B"]), int(g["N"]), int(g["seed"])
    visual, question, adj, n_obj = syn.make_explicit_inputs(cfg["v_dim"], cfg["q_dim"], cfg["label_num"], B, N, seed)
    np.testing.assert_allclose([visual.sum(), question.sum(), adj.sum()], g["input_check"], rtol=1e-12)
    shapes = [ast.literal_eval(s) for s in g["shapes"]]
    params = syn.explicit_param_values(shapes, seed + 1)
    np.testing.assert_allclose([p.sum() for p in params], g["param_check"], rtol=1e-10, atol=1e-12)
    return g, cfg, B, N, visual, question, adj, [str(n) for n in g["names"]], params


def test_fixtures_exist():
    assert len(FILES) >= 5 and any(os.path.getsize(f) for f in FILES)
    steps = sorted(ast.literal_eval(str(np.load(f)["cfg"])).get("num_steps", 1) for f in FILES)
    assert steps[-2:] == [2, 3]                       # the encoder's num_steps argument is covered beyond the default


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_matches_reference_execution(path):
    from oracle import explicit_relation as oe
    g, cfg, B, N, visual, question, adj, names, params = _case(path)
    p = {n: torch.tensor(a, dtype=torch.float64, requires_grad=True) for n, a in zip(names, params)}
    tv, tq = torch.tensor(visual, requires_grad=True), torch.tensor(question, requires_grad=True)
    out = oe.forward(p, cfg, tv, torch.tensor(adj), tq)
    assert np.abs(out.detach().numpy() - g["output"]).max() < 1e-5 * np.abs(g["output"]).max()      # stored as float32
    probe = torch.tensor(np.random.default_rng(77).standard_normal((B, N, cfg["out_dim"])))
    grads = torch.autograd.grad((out * probe).sum(), list(p.values()) + [tv, tq], allow_unused=True)
    for i, (n, gr) in enumerate(zip(names, grads[:-2])):
        a = np.zeros(p[n].shape) if gr is None else gr.numpy()
        assert abs(np.sqrt((a * a).sum()) - float(g["grad.norm/" + n])) <= 1e-9 * max(float(g["grad.norm/" + n]), 1e-12) + 1e-12, n
        np.testing.assert_allclose(a.ravel()[g["grad.idx/" + n]], g["grad.sample/" + n], rtol=1e-8, atol=1e-11, err_msg=n)
    assert np.abs(grads[-2].numpy() - g["grad_visual"]).max() < 1e-5 * np.abs(g["grad_visual"]).max() + 1e-9
    assert np.abs(grads[-1].numpy() - g["grad_question"]).max() < 1e-5 * np.abs(g["grad_question"]).max() + 1e-9


def test_build_graph_matches_reference_function():
    from tf_vqa_regat_b200.model.position_emb import build_graph, one_hot_adjacency
    g = np.load(os.path.join(HERE, "golden", "explicit_build_graph.npz"))
    for k in range(3):
        got = build_graph(g[f"bbox{k}"], g[f"spatial{k}"])
        np.testing.assert_array_equal(got, g[f"adj{k}"])            # integer labels: bit-exact
    lab = g["adj2"]
    oh = one_hot_adjacency(lab, 11)
    assert oh.shape == lab.shape + (11,) and oh.sum() == ((lab > 0) & (lab <= 11)).sum()
    i, j = np.argwhere((lab > 0) & (lab <= 11))[0]
    assert oh[i, j, int(lab[i, j]) - 1] == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_layer_mirror_matches_reference_execution(path):
    from tf_vqa_regat_b200.model import ExplicitRelationEncoder
    g, cfg, B, N, visual, question, adj, names, params = _case(path)
    enc = ExplicitRelationEncoder(cfg["v_dim"], cfg["q_dim"], cfg["out_dim"], cfg["dir_num"], cfg["label_num"], nongt_dim=cfg["nongt_dim"],
                                  num_heads=cfg["num_heads"], num_steps=cfg.get("num_steps", 1), residual_connection=cfg["residual"],
                                  label_bias=cfg["label_bias"])       # *_steps2 / *_steps3 fixtures: relation_encoder.py:134-141
    f32 = lambda a: torch.tensor(np.asarray(a, dtype=np.float32)).cuda()
    enc(f32(visual), f32(adj), f32(question))                        # creates the variables
    got_names = [n for n, _ in enc.weights]
    assert len(got_names) == len(names) and [n.split("/")[-1] for n in names] == got_names          # v, g, bias per layer, same order
    enc.set_weights([np.asarray(p, dtype=np.float32) for p in params])
    out = enc(f32(visual), f32(adj), f32(question)).cpu().numpy()
    ref = g["output"]
    assert np.abs(out - ref).max() < 1e-4 * np.abs(ref).max()
    # both spellings of the constructor argument, and the shape check
    ExplicitRelationEncoder(8, 8, 8, 1, 11, residiual_connection=False)
    with pytest.raises(ValueError):
        enc(f32(visual), f32(adj[..., :3]), f32(question))


def _torch_attention(q, kv, pb, s, H, dirs, M):
    B, N, _ = s.shape
    D = s.shape[-1]
    dh = D // H
    total = s
    ps = []
    for d in range(dirs):
        Q = q[..., d * D:(d + 1) * D].view(B, N, H, dh).transpose(1, 2)
        K = kv[..., d * D:(d + 1) * D].view(B, M, H, dh).transpose(1, 2)
        V = kv[..., (dirs + d) * D:(dirs + d + 1) * D].view(B, M, H, dh).transpose(1, 2)
        aff = Q @ K.transpose(-1, -2) / dh ** 0.5
        b = pb[:, d][:, None]
        logits = torch.where(b > -1e15, aff + b, b.expand_as(aff))
        P = torch.softmax(logits, -1)
        ps.append(P)
        total = total + (P @ V).transpose(1, 2).reshape(B, N, D)
    return torch.relu(total), ps


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_masked_attention_kernels_vs_torch(dtype, tol):
    """regat_explicit_pair_bias(+_bwd), regat_graphattn_explicit_fwd / _bwd on random operands with fully masked rows (padded
    objects): output, dQ, dK, dV', ds and the label FC's gradients against autograd of the plain formulation."""
    import ctypes as C
    from tf_vqa_regat_b200 import _lib
    l = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(20260318)      # operands must not depend on what ran before (the bf16 bound is a max-norm over random data)
    B, N, nongt, D, H, dirs, L = 3, 36, 20, 256, 4, 2, 11
    M = min(nongt, N)
    rng = np.random.default_rng(3)
    _, _, adj, _ = syn.make_explicit_inputs(8, 8, L, B, N, seed=11)
    adj_t = torch.tensor(adj, dtype=torch.float32).cuda()
    w = torch.tensor(rng.standard_normal(L), dtype=torch.float32, device="cuda", requires_grad=True)
    bl = torch.tensor([0.3], dtype=torch.float32, device="cuda", requires_grad=True)
    code, tdt = (_lib.F32, torch.float32) if dtype == "fp32" else (_lib.BF16, torch.bfloat16)
    mk = lambda *s_: (0.5 * torch.randn(*s_, device="cuda")).to(tdt).float().requires_grad_(True)       # values exact in the kernel dtype
    q, kv, s = mk(B, N, dirs * D), mk(B, M, 2 * dirs * D), mk(B, N, D)
    # reference
    pbs = []
    for d in range(dirs):
        a = adj_t if d == 0 else adj_t.transpose(1, 2)
        a = a[:, :, :M]
        lab = a @ w + bl
        pbs.append(torch.where(a.sum(-1) > 0, lab, torch.full_like(lab, -9e15)))
    pb_ref = torch.stack(pbs, 1)
    out_ref, P_ref = _torch_attention(q, kv, pb_ref, s, H, dirs, M)
    probe = torch.randn(B, N, D, device="cuda").to(tdt).float()
    gq, gkv, gs, gw, gb = torch.autograd.grad((out_ref * probe).sum(), [q, kv, s, w, bl])
    # kernels
    pb = torch.empty(B, dirs, N, M, device="cuda")
    _lib.check(l.regat_explicit_pair_bias(B, N, nongt, L, dirs, adj_t.data_ptr(), w.data_ptr(), bl.data_ptr(), pb.data_ptr(), st))
    assert torch.equal(pb > -1e15, pb_ref > -1e15) and torch.allclose(pb, pb_ref.detach(), rtol=1e-6, atol=1e-6)
    cast = lambda t: t.detach().to(tdt).contiguous()
    qd, kvd, sd = cast(q), cast(kv), cast(s)
    v1 = torch.empty(B, N, D, device="cuda", dtype=tdt)
    P = torch.zeros(B, dirs, H, N, M, device="cuda")
    gate = torch.zeros(B * N * H, dtype=torch.int64, device="cuda")
    _lib.check(l.regat_graphattn_explicit_fwd(code, B, N, nongt, D, H, dirs, qd.data_ptr(), kvd.data_ptr(), pb.data_ptr(), sd.data_ptr(), None, 0,
                                              v1.data_ptr(), P.data_ptr(), gate.data_ptr(), st))
    rel = lambda a, b: float((a.float() - b).abs().max() / b.abs().max())
    assert rel(v1, out_ref.detach()) < tol
    assert rel(P, torch.stack(P_ref, 1).detach()) < tol
    dq, dkv, dout = torch.empty_like(qd), torch.empty_like(kvd), torch.empty_like(sd)
    _lib.check(l.regat_graphattn_explicit_bwd(code, B, N, nongt, D, H, dirs, qd.data_ptr(), kvd.data_ptr(), cast(probe).data_ptr(), gate.data_ptr(),
                                              pb.data_ptr(), P.data_ptr(), dq.data_ptr(), dkv.data_ptr(), dout.data_ptr(), st))
    if dtype == "fp32":
        assert rel(dq, gq) < tol and rel(dkv, gkv) < tol and rel(dout, gs) < tol, (rel(dq, gq), rel(dkv, gkv), rel(dout, gs))
    else:
        # bf16: a relu pre-activation within rounding distance of 0 lands on the other side in bf16 and its element of the gated
        # gradient appears or disappears whole, so single elements are bounded loosely (0.3 of the largest entry) and the tensors
        # in the 2-norm (bf16 storage of dQ / dK / dV' on top of single-pass TF32 products: measured 1e-2)
        fro = lambda a, b: float((a.float() - b).norm() / b.norm())
        errs = (fro(dq, gq), fro(dkv, gkv), fro(dout, gs), rel(dq, gq), rel(dkv, gkv), rel(dout, gs))
        assert max(errs[:3]) < 4e-2 and max(errs[3:]) < 0.3, errs
    dw, db = torch.zeros(L, device="cuda"), torch.zeros(1, device="cuda")
    _lib.check(l.regat_explicit_pair_bias_bwd(B, N, nongt, L, dirs, H, adj_t.data_ptr(), P.data_ptr(), dw.data_ptr(), db.data_ptr(), st))
    torch.cuda.synchronize()
    scale = float(gw.abs().max())
    assert float((dw - gw).abs().max()) < (tol if dtype == "fp32" else 8e-2) * scale + 1e-5
    # db is a zero direction (softmax shift invariance within live rows; fully masked rows are uniform): rounding noise only
    assert abs(float(db) - float(gb)) < 1e-2 * scale + 1e-4
