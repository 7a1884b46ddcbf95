"""Host logic of tf_vqa_regat_b200/question.py (the launch sequence of the question front-end) dry-run on the
CPU: the module runs unchanged, with the library's entry points replaced by tests/_host_emulation.py, and must reproduce
the oracle that is pinned to the executed reference (tests/golden/refexec_question_*.npz).  What this covers: operand
shapes / transposes / leading dimensions of every GEMM, buffer offsets, the back-propagation through time, weight-norm
backward, clip + Adamax.  What it cannot cover -- the CUDA kernels themselves -- is tests/test_gpu_question.py."""
import numpy as np
import pytest
import torch

from oracle import language_model as olm
from oracle import regat_torch as ot
from tf_vqa_regat_b200.question import QuestionFrontEnd, question_layout

from _host_emulation import HostOps


def _setup(op, B, emb2_trainable, n_token=40, E=10, H=24, T=14, max_batch=None):
    fe = QuestionFrontEnd(n_token, E, H, op=op, seq_len=T, max_batch=max_batch or B, emb2_trainable=emb2_trainable, device="cpu",
                          _ops=HostOps())
    front = olm.make_params(n_token, E, H, op, seed=11)
    fe.load_named(front)
    tok = olm.make_tokens(B, n_token, T, seed=21)
    return fe, front, tok


def _oracle(front, tok, n_token, op, dq_att, dq_last, trainable):
    p = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in front.items()}
    q = olm.forward(p, tok, n_token, op)
    q["w_emb"].retain_grad()
    loss = (q["q_att"] * torch.tensor(dq_att, dtype=torch.float64)).sum() + (q["q_last"] * torch.tensor(dq_last, dtype=torch.float64)).sum()
    loss.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in p.items()}
    grads["__w_emb_occurrences"] = q["w_emb"].grad.numpy()          # IndexedSlices.values of the tables (before the padding mask)
    return q, grads


def test_layout_matches_the_oracle_order():
    ent, total = question_layout(40, 10, 24, "c")
    assert [(n, tuple(s)) for n, s, _ in ent] == [(n, tuple(s)) for n, s, _ in olm.param_shapes(40, 10, 24, "c")]
    assert all(o % 64 == 0 for _, _, o in ent) and total % 64 == 0
    assert [n for n, _, _ in question_layout(40, 10, 24, "")[0]][1] == "q_emb.gru/kernel"


@pytest.mark.parametrize("op,B,emb2_tr,max_batch", [("c", 3, False, None), ("c", 4, True, 6), ("", 2, False, None)])
def test_forward_backward_sequence_matches_oracle(op, B, emb2_tr, max_batch):
    fe, front, tok = _setup(op, B, emb2_tr, max_batch=max_batch)
    q_att, q_last = fe.forward(torch.tensor(tok, dtype=torch.int32))
    rng = np.random.default_rng(5)
    dqa, dql = rng.standard_normal((B, 24)).astype(np.float32), rng.standard_normal((B, 24)).astype(np.float32)
    q, grads = _oracle(front, tok, 40, op, dqa, dql, None)
    np.testing.assert_allclose(q_att.numpy(), q["q_att"].detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(q_last.numpy(), q["q_last"].detach().numpy(), rtol=2e-5, atol=2e-6)
    fe.backward(torch.tensor(dqa), torch.tensor(dql))
    got = {k: v.numpy() for k, v in fe.named(fe.grads).items()}
    for name in got:
        if name == "w_emb.emb_/emb_" and not emb2_tr:
            assert not got[name].any()                                   # frozen table: no gradient is produced
            continue
        scale = max(np.abs(grads[name]).max(), 1e-6)
        assert np.abs(got[name] - grads[name]).max() < 5e-5 * scale + 1e-7, name
    pad_rows = got["w_emb.emb/emb"][40]
    assert not pad_rows.any()                                             # the padding row never receives gradient


def test_update_is_per_tensor_clip_then_adamax():
    fe, front, tok = _setup("c", 3, True)
    rng = np.random.default_rng(6)
    dqa, dql = rng.standard_normal((3, 24)).astype(np.float32), rng.standard_normal((3, 24)).astype(np.float32)
    p = {k: np.asarray(v, dtype=np.float64) for k, v in front.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}; u = {k: np.zeros_like(v) for k, v in p.items()}
    for step in (1, 2, 3):                                                # m / u carry over between steps
        fe.forward(torch.tensor(tok, dtype=torch.int32))
        fe.backward(torch.tensor(dqa), torch.tensor(dql))
        fe.update(1e-3, step)
        _, grads = _oracle(p, tok, 40, "c", dqa, dql, None)
        occ = grads["__w_emb_occurrences"]
        for k in p:
            if k.startswith("w_emb."):         # tf.IndexedSlices semantics: occurrence-norm clip, sparse Adamax (train.py:112-113)
                vals = occ[..., :10] if k == "w_emb.emb/emb" else occ[..., 10:]
                p[k], m[k], u[k] = olm.sparse_clip_adamax(p[k], tok, vals, m[k], u[k], step, 1e-3, 0.25, 0.9, 0.999, 1e-8, 40)
                continue
            p[k], m[k], u[k] = ot.adamax_step(p[k], ot.clip_by_norm(grads[k], 0.25), m[k], u[k], step, 1e-3)
    got = {k: v.numpy() for k, v in fe.named().items()}
    for k in p:
        if k == "q_att.linear2/bias":
            continue     # a constant added to every logit of a position: softmax-shift invariant, gradient is rounding noise -> +-lr
        # an Adamax step is at most lr per element; fp32 storage of the parameters and of m / u is the only difference
        assert np.abs(got[k] - p[k]).max() < 0.02 * 3 * 1e-3, k
    assert np.abs(got["w_emb.emb/emb"] - np.asarray(front["w_emb.emb/emb"])).max() > 1e-3      # it did move


def test_input_validation():
    fe, _, tok = _setup("c", 3, False)
    with pytest.raises(ValueError):
        fe.forward(torch.tensor(tok[:1], dtype=torch.int32))             # batch 1 (language_model.py:159 squeeze)
    with pytest.raises(ValueError):
        fe.forward(torch.tensor(tok, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        fe.backward(torch.zeros(3, 24), torch.zeros(3, 24))
    from tf_vqa_regat_b200._lib import RegatError
    if not torch.cuda.is_available():
        with pytest.raises(RegatError):
            QuestionFrontEnd(40, 10, 24)                                 # no CUDA, no emulation handed in: refuses


def _dp_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    n_token, E, H, T, Bg = 40, 10, 24, 14, 6
    front = olm.make_params(n_token, E, H, "c", seed=11)
    tok = olm.make_tokens(Bg, n_token, T, seed=21)
    rng = np.random.default_rng(5)
    dqa, dql = rng.standard_normal((Bg, H)).astype(np.float32), rng.standard_normal((Bg, H)).astype(np.float32)
    Bl = Bg // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    fe = QuestionFrontEnd(n_token, E, H, op="c", seq_len=T, max_batch=Bl, emb2_trainable=True, device="cpu", group=dist.group.WORLD,
                          _ops=HostOps())
    fe.load_named(front)
    q_att, q_last = fe.forward(torch.tensor(tok[sl], dtype=torch.int32))
    fe.backward(torch.tensor(dqa[sl]), torch.tensor(dql[sl]))
    fe.allreduce_grads()
    grads = {k: v.numpy().copy() for k, v in fe.named(fe.grads).items()}
    fe.update(1e-3, 1)                      # embedding tables: occurrence-norm clip + sparse Adamax over the GLOBAL batch's tokens
    out[rank] = (q_att.numpy().copy(), q_last.numpy().copy(), grads, {k: v.numpy().copy() for k, v in fe.named().items()})
    dist.destroy_process_group()


def test_two_rank_gloo_front_end_reproduces_the_global_batch():
    """SURVEY 8e for the front-end: its softmax runs over the batch axis (language_model.py:163-167), so sharding needs an
    exchange -- the ranks gather each other's [B_local, T] logits (and dW rows in the backward pass).  Two gloo ranks on halves
    of a batch of 6 must give the single-process oracle's q_att / q_last rows and, after summing over ranks, its gradients."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, 29523, out), nprocs=2, join=True)
    n_token, E, H, T, Bg = 40, 10, 24, 14, 6
    front = olm.make_params(n_token, E, H, "c", seed=11)
    tok = olm.make_tokens(Bg, n_token, T, seed=21)
    rng = np.random.default_rng(5)
    dqa, dql = rng.standard_normal((Bg, H)).astype(np.float32), rng.standard_normal((Bg, H)).astype(np.float32)
    q, grads = _oracle(front, tok, n_token, "c", dqa, dql, None)
    for r in range(2):
        sl = slice(r * 3, (r + 1) * 3)
        np.testing.assert_allclose(out[r][0], q["q_att"].detach().numpy()[sl], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(out[r][1], q["q_last"].detach().numpy()[sl], rtol=2e-5, atol=2e-6)
        for name, g in out[r][2].items():
            if name == "q_att.linear2/bias":
                assert np.abs(g).max() < 1e-5          # softmax-shift direction: exactly zero in real arithmetic, rounding noise here
                continue
            scale = max(np.abs(grads[name]).max(), 1e-6)
            assert np.abs(g - grads[name]).max() < 5e-5 * scale + 1e-7, (r, name)
    # one optimizer step: both ranks hold the single-process result (the embedding tables need every rank's token occurrences)
    occ = grads["__w_emb_occurrences"]
    for name in ("w_emb.emb/emb", "w_emb.emb_/emb_"):
        vals = occ[..., :E] if name == "w_emb.emb/emb" else occ[..., E:]
        z = np.zeros_like(np.asarray(front[name], dtype=np.float64))
        want, _, _ = olm.sparse_clip_adamax(np.asarray(front[name], dtype=np.float64), tok, vals, z, z.copy(), 1, 1e-3, 0.25, 0.9, 0.999, 1e-8, n_token)
        for r in range(2):
            assert np.abs(out[r][3][name] - want).max() < 0.02 * 1e-3, (r, name)
        assert np.array_equal(out[0][3][name], out[1][3][name])
    # and it matters: a rank that normalised over its own half only would be far off
    p = {k: torch.tensor(np.asarray(v, dtype=np.float64)) for k, v in front.items()}
    local_only = olm.forward(p, tok[:3], n_token, "c")["q_att"].numpy()
    assert np.abs(local_only - q["q_att"].detach().numpy()[:3]).max() > 1e-2
