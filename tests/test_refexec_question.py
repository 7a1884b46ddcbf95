"""Question front-end and the whole model, tokens -> logits: the oracle (oracle/language_model.py + the hot-path
oracle) against vectors produced by executing the reference's own language_model.py / rel_graph_net.py / train.py over
oracle/tf_shim (oracle/make_golden_ref_question.py)."""
import ast
import glob
import os
import sys

import numpy as np
import pytest
import torch

from oracle import language_model as olm
from oracle import regat_torch as ot
from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig, param_layout

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "refexec_question_*.npz")))
IDS = [os.path.basename(f)[len("refexec_"):-4] for f in FILES]


def _load(path):
    g = np.load(path)
    cfg = HotPathConfig(**ast.literal_eval(str(g["cfg"])))
    B, N, steps = int(g["B"]), int(g["N"]), int(g["steps"])
    n_token, emb_dim, op = int(g["n_token"]), int(g["emb_dim"]), str(g["op"])
    batches = [syn.make_inputs(cfg, B, N, seed=1000 + s, adaptive=bool(g["adaptive"])) for s in range(steps + 1)]
    np.testing.assert_allclose([float(np.sum(b["features"], dtype=np.float64)) for b in batches], g["input_check"], rtol=1e-12)
    tokens = [olm.make_tokens(B, n_token, 14, seed=21 + s) for s in range(steps + 1)]
    assert np.array_equal(np.stack(tokens), g["tokens"])
    hot = syn.unflatten(cfg, syn.make_params(cfg, seed=7, trained_like=True).astype(np.float64))
    front = {k: np.asarray(v, dtype=np.float64) for k, v in olm.make_params(n_token, emb_dim, cfg.q_dim, op, seed=11).items()}
    shapes = olm.param_shapes(n_token, emb_dim, cfg.q_dim, op, bool(g["emb2_trainable"]))
    return g, cfg, batches, tokens, hot, front, shapes, n_token, op


def _loss_and_grads(cfg, hot, front, shapes, batch, tok, n_token, op):
    trainable = {n for n, _, t in shapes if t}
    pf = {k: torch.tensor(v, dtype=torch.float64, requires_grad=k in trainable) for k, v in front.items()}
    ph = ot.to_torch_params(hot)
    q = olm.forward(pf, tok, n_token, op)
    q["w_emb"].retain_grad()
    out = ot.forward(ph, cfg, batch["features"], batch["boxes"], q["q_att"], q["q_last"], batch["target"])
    out["loss"].backward()
    grads = {k: v.grad.numpy() for k, v in list(pf.items()) + list(ph.items()) if v.grad is not None}
    # per-occurrence gradient of the (masked) lookup output: what TensorFlow hands over as IndexedSlices.values for the tables
    grads["__w_emb_occurrences"] = q["w_emb"].grad.numpy()
    return q, out, grads


def test_fixtures_present():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_front_end_and_whole_model_match_reference_execution(path):
    g, cfg, batches, tokens, hot, front, shapes, n_token, op = _load(path)
    q, out, grads = _loss_and_grads(cfg, hot, front, shapes, batches[0], tokens[0], n_token, op)
    for k in ("w_emb", "q_seq", "q_att", "q_last"):
        np.testing.assert_allclose(q[k].detach().numpy(), g[k], rtol=1e-10, atol=1e-13, err_msg=k)
    pad = tokens[0] == n_token
    assert pad.any() and not np.abs(g["w_emb"][pad]).any()                # padding rows are zeroed (language_model.py:33-38)
    assert np.abs(g["q_seq"][pad]).min() > 0                              # ... but still run through the GRU (no masking)
    np.testing.assert_allclose(out["logits"].detach().numpy(), g["logits"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(float(out["loss"].detach()), float(g["loss"]), rtol=1e-12)
    for n, _, t in shapes:
        if t:
            np.testing.assert_allclose(grads[n], g["grad/" + n], rtol=1e-8, atol=1e-13, err_msg=n)
        else:
            assert "grad/" + n not in g.files                              # frozen second table (language_model.py:58)
    for e in param_layout(cfg)[0]:
        np.testing.assert_allclose(np.sqrt((grads[e.name] ** 2).sum()), float(g["gradnorm/" + e.name]), rtol=1e-8, atol=1e-13)


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_batch_axis_softmax_is_what_the_reference_does(path):
    """language_model.py:163-167: the attention weights of the question are normalised over the batch and raw-reshaped.  The
    'sensible' per-question softmax gives a different vector; the executed reference agrees with the quirk, not with it."""
    g, cfg, batches, tokens, hot, front, shapes, n_token, op = _load(path)
    pf = {k: torch.tensor(v) for k, v in front.items()}
    seq = torch.tensor(g["q_seq"])
    quirk = olm.question_self_attention(pf, seq).numpy()
    a1 = torch.tanh(olm._fc(seq, pf, "q_att.linear1"))
    w = torch.softmax(olm._fc(a1, pf, "q_att.linear2").squeeze(-1), dim=1)          # per question, over positions
    sensible = torch.matmul(w.unsqueeze(1), seq).squeeze(1).numpy()
    np.testing.assert_allclose(quirk, g["q_att"], rtol=1e-10, atol=1e-13)
    assert np.abs(sensible - g["q_att"]).max() > 1e-2


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_train_loop_updates_front_end_like_reference_train(path):
    g, cfg, batches, tokens, hot, front, shapes, n_token, op = _load(path)
    steps, lr = int(g["steps"]), float(g["lr"])
    p = {**{k: v.copy() for k, v in front.items()}, **{k: v.copy() for k, v in hot.items()}}
    trainable = [n for n, _, t in shapes if t] + [e.name for e in param_layout(cfg)[0]]
    m = {k: np.zeros_like(p[k]) for k in trainable}
    u = {k: np.zeros_like(p[k]) for k in trainable}
    for step in range(1, steps + 1):
        fr = {k: p[k] for k in front}
        ho = {k: p[k] for k in hot}
        _, out, grads = _loss_and_grads(cfg, ho, fr, shapes, batches[step - 1], tokens[step - 1], n_token, op)
        np.testing.assert_allclose(out["logits"].detach().numpy(), g["train.logits"][step - 1], rtol=1e-8, atol=1e-10)
        occ = grads["__w_emb_occurrences"]
        E = p["w_emb.emb/emb"].shape[1]
        for k in trainable:
            if k.startswith("w_emb."):          # IndexedSlices: clipped by the occurrence norm, sparse Adamax (oracle/language_model.py)
                vals = occ[..., :E] if k == "w_emb.emb/emb" else occ[..., E:]
                p[k], m[k], u[k] = olm.sparse_clip_adamax(p[k], tokens[step - 1], vals, m[k], u[k], step, lr, cfg.grad_clip, cfg.beta1, cfg.beta2,
                                                          cfg.eps, n_token)
                continue
            gk = ot.clip_by_norm(grads[k], cfg.grad_clip)
            p[k], m[k], u[k] = ot.adamax_step(p[k], gk, m[k], u[k], step, lr, cfg.beta1, cfg.beta2, cfg.eps)
    for n, _, t in shapes:
        np.testing.assert_allclose(p[n], g["param/" + n], rtol=0, atol=1e-9 * max(np.abs(p[n]).max(), 1.0), err_msg=n)
        if not t:
            assert np.array_equal(p[n], front[n])
    # rows of the embedding table that no token touched do not move: gradient exactly 0 -> m = u = 0 -> 0 / (0 + eps)
    used = np.unique(np.stack(tokens[:steps]))
    untouched = np.setdiff1d(np.arange(n_token + 1), used)
    assert np.array_equal(g["param/w_emb.emb/emb"][untouched], front["w_emb.emb/emb"][untouched])


def test_shim_gru_matches_torch_gru():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "tf_shim"))
    try:
        import tensorflow as tf
        tf.keras.backend.set_floatx("float64")
        rng = np.random.default_rng(4)
        gru = tf.keras.layers.GRU(units=7, return_sequences=True, return_state=True)
        x = rng.standard_normal((3, 5, 4))
        seq, last = gru(x)
        gru.bias.assign(0.3 * rng.standard_normal((2, 21)))
        seq, last = gru(x)
        assert [w.var_name for w in gru.weights] == ["kernel", "recurrent_kernel", "bias"]
        u = 7
        perm = lambda m: np.concatenate([m[:, u:2 * u], m[:, :u], m[:, 2 * u:]], 1)       # keras z|r|h -> torch r|z|n
        tg = torch.nn.GRU(4, 7, batch_first=True).double()
        with torch.no_grad():
            tg.weight_ih_l0.copy_(torch.tensor(perm(gru.kernel.numpy()).T)); tg.weight_hh_l0.copy_(torch.tensor(perm(gru.recurrent_kernel.numpy()).T))
            tg.bias_ih_l0.copy_(torch.tensor(perm(gru.bias.numpy()[0:1])[0])); tg.bias_hh_l0.copy_(torch.tensor(perm(gru.bias.numpy()[1:2])[0]))
        want, hn = tg(torch.tensor(x))
        np.testing.assert_allclose(seq.numpy(), want.detach().numpy(), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(last.numpy(), hn[0].detach().numpy(), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(olm.gru({"q_emb.gru/kernel": torch.tensor(gru.kernel.numpy()), "q_emb.gru/recurrent_kernel":
                                            torch.tensor(gru.recurrent_kernel.numpy()), "q_emb.gru/bias": torch.tensor(gru.bias.numpy())},
                                           torch.tensor(x)).numpy(), want.detach().numpy(), rtol=1e-12, atol=1e-14)
    finally:
        sys.path.remove(os.path.join(os.path.dirname(HERE), "oracle", "tf_shim"))
        for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
            del sys.modules[k]
