"""The oracle against vectors produced by EXECUTING the reference's own model/*.py and train.py
(oracle/make_golden_ref.py; TensorFlow's primitives supplied by oracle/tf_shim).  These are the
vectors that pin stages 2-3, the loss, clip_by_norm and Adamax; tests/golden/hotpath_*.npz are
only regression anchors.  Also unit-checks the stand-in's primitives against independent
implementations, and -- in the build container, where /root/reference exists -- re-runs the
reference and compares with the committed files."""
import ast
import glob
import os
import sys

import numpy as np
import pytest
import torch

from oracle import regat_fused as ofu
from oracle import regat_numpy as onp
from oracle import regat_torch as ot
from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig, param_layout

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = [f for f in sorted(glob.glob(os.path.join(HERE, "golden", "refexec_*.npz"))) if "refexec_question_" not in f and "refexec_collate" not in f
         and "refexec_explicit_" not in f]      # explicit relation encoders: tests/test_explicit.py
IDS = [os.path.basename(f)[len("refexec_"):-4] for f in FILES]


def load_case(path):
    g = np.load(path)
    cfg = HotPathConfig(**ast.literal_eval(str(g["cfg"])))
    B, N, steps = int(g["B"]), int(g["N"]), int(g["steps"])
    batches = [syn.make_inputs(cfg, B, N, seed=1000 + s, adaptive=bool(g["adaptive"])) for s in range(steps + 1)]
    # the fixtures were made from these very inputs
    np.testing.assert_allclose([float(np.sum(b["features"], dtype=np.float64)) for b in batches], g["input_check"], rtol=1e-12)
    flat = syn.make_params(cfg, seed=7, trained_like=bool(g["trained_like"]))
    return g, cfg, batches, flat


def check_summary(g, prefix, name, a, seed, rtol, atol):
    """Compare a tensor with the (sum, norm, seeded projection, sample[, full]) record of make_golden_ref._summ."""
    a = np.asarray(a, dtype=np.float64)
    r = np.random.default_rng(seed).standard_normal(a.size)
    scale = float(g[f"{prefix}.norm/{name}"])
    np.testing.assert_allclose(np.sqrt((a * a).sum()), scale, rtol=rtol, atol=atol, err_msg=name)
    np.testing.assert_allclose(a.ravel() @ r, float(g[f"{prefix}.proj/{name}"]), rtol=0, atol=rtol * scale * 4 + atol, err_msg=name)
    np.testing.assert_allclose(a.ravel()[g[f"{prefix}.idx/{name}"]], g[f"{prefix}.sample/{name}"], rtol=0,
                               atol=rtol * max(np.abs(a).max(), 1e-300) + atol, err_msg=name)
    if f"{prefix}.full/{name}" in g:
        np.testing.assert_allclose(a, g[f"{prefix}.full/{name}"], rtol=0, atol=rtol * max(np.abs(a).max(), 1e-300) + atol, err_msg=name)


def test_fixtures_present():
    assert len(FILES) >= 10


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_forward_and_gradients_match_reference_execution(path):
    g, cfg, batches, flat = load_case(path)
    inp = batches[0]
    named = syn.unflatten(cfg, flat.astype(np.float64))
    loss, grads, dq_att, dq_last, out = ot.loss_and_grads(named, cfg, inp)
    np.testing.assert_allclose(loss, float(g["loss"]), rtol=1e-12)
    for k in ("logits", "joint", "att_weights"):
        np.testing.assert_allclose(out[k], g[k], rtol=1e-10, atol=1e-12, err_msg=k)
    if "v1" in g:
        np.testing.assert_allclose(out["v1"], g["v1"], rtol=1e-10, atol=1e-12)
    else:
        np.testing.assert_allclose(out["v1"][:, :, :16], g["v1_head"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out["v1"].sum(-1), g["v1_sum"], rtol=1e-10)
    np.testing.assert_allclose(dq_att, g["dq_att"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(dq_last, g["dq_last"], rtol=1e-9, atol=1e-13)
    for i, e in enumerate(param_layout(cfg)[0]):
        check_summary(g, "grad", e.name, grads[e.name], 100 + i, rtol=1e-9, atol=1e-13)
    # the NumPy transcription and the re-associated (kernel) formulation see the same vectors
    f64 = lambda a: a.astype(np.float64)
    args = (f64(inp["features"]), inp["boxes"], f64(inp["q_att"]), f64(inp["q_last"]), f64(inp["target"]))
    a = onp.forward(named, cfg, *args)
    np.testing.assert_allclose(a["logits"], g["logits"], rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(a["mask"], g["mask"])                      # q-mask: exact
    c = ofu.forward(named, cfg, *args)
    np.testing.assert_allclose(c["logits"], g["logits"], rtol=1e-8, atol=1e-10)
    np.testing.assert_array_equal(c["mask"], g["mask"])
    assert np.array_equal(a["logits"].argmax(1), g["logits"].argmax(1))


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_fp32_noise_floor_of_the_reference(path):
    """What the reference's own float32 arithmetic does to the logits: the 1e-4 parity budget sits two orders above it."""
    g = np.load(path)
    err = np.abs(g["logits_f32"] - g["logits"]).max() / np.abs(g["logits"]).max()
    assert err < 2e-6, err


@pytest.mark.parametrize("path", [f for f in FILES if "full_b4" not in f], ids=[i for i in IDS if "full_b4" not in i])
def test_train_loop_matches_reference_train(path):
    """train.train(): GradientTape -> per-tensor clip_by_norm(0.25) -> Adamax, `steps` batches; then train.evaluate()."""
    g, cfg, batches, flat = load_case(path)
    steps, lr = int(g["steps"]), float(g["lr"])
    p = {k: v.copy() for k, v in syn.unflatten(cfg, flat.astype(np.float64)).items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    u = {k: np.zeros_like(v) for k, v in p.items()}
    for step in range(1, steps + 1):
        loss, grads, _, _, out = ot.loss_and_grads(p, cfg, batches[step - 1])
        np.testing.assert_allclose(out["logits"], g["train.logits"][step - 1], rtol=1e-9, atol=1e-11)
        for k in p:
            gk = ot.clip_by_norm(grads[k], cfg.grad_clip)
            p[k], m[k], u[k] = ot.adamax_step(p[k], gk, m[k], u[k], step, lr, cfg.beta1, cfg.beta2, cfg.eps)
    for i, e in enumerate(param_layout(cfg)[0]):
        zero_dir = ("implicit_relation.bias/" in e.name or e.name.endswith(".key/bias")
                    or e.name in ("joint_emb.linear/bias", "joint_emb.v2attention/bias"))
        if zero_dir:
            # exactly-zero directions: the gradient is rounding noise, Adamax turns it into +-lr steps
            assert np.abs(p[e.name] - syn.unflatten(cfg, flat.astype(np.float64))[e.name]).max() <= steps * lr * 1.0001
            continue
        check_summary(g, "param", e.name, p[e.name], 500 + i, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(np.sqrt((m[e.name] ** 2).sum()), float(g[f"adamax_m.norm/{e.name}"]), rtol=1e-8, atol=1e-15)
        np.testing.assert_allclose(np.sqrt((u[e.name] ** 2).sum()), float(g[f"adamax_u.norm/{e.name}"]), rtol=1e-8, atol=1e-15)
    ev = batches[steps]
    f64 = lambda a: a.astype(np.float64)
    lo = onp.forward(p, cfg, f64(ev["features"]), ev["boxes"], f64(ev["q_att"]), f64(ev["q_last"]), f64(ev["target"]))["logits"]
    np.testing.assert_allclose(lo, g["eval.logits"], rtol=1e-7, atol=1e-9)
    score = 100.0 * float(np.take_along_axis(ev["target"], lo.argmax(1)[:, None], 1).sum()) / ev["target"].shape[0]
    assert abs(score - float(g["eval.score_pct"])) < 1e-3                    # train.py:28-39,171-172 (printed with 4 decimals)


# ------------------------------------------------------------------ the stand-in's primitives, checked independently
@pytest.fixture(scope="module")
def tf():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "tf_shim"))
    try:
        import tensorflow as tf
        assert tf.__version__.endswith("standin")
        tf.keras.backend.set_floatx("float64")
        yield tf
    finally:
        sys.path.remove(os.path.join(os.path.dirname(HERE), "oracle", "tf_shim"))
        for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
            del sys.modules[k]                        # nothing else in this process may mistake the stand-in for TensorFlow


def test_shim_grouped_conv_matches_torch_conv2d(tf):
    rng = np.random.default_rng(0)
    conv = tf.keras.layers.Conv2D(filters=12, kernel_size=(1, 1), groups=4)
    x = rng.standard_normal((5, 1, 1, 32))
    y = conv(x).numpy()
    k = conv.kernel.numpy()                                             # [1,1,Cin/G,Cout]
    want = torch.nn.functional.conv2d(torch.tensor(x).permute(0, 3, 1, 2), torch.tensor(k).permute(3, 2, 0, 1),
                                      torch.tensor(conv.bias.numpy()), groups=4).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(y, want, rtol=1e-12, atol=1e-14)
    # and against the definition, group by group
    for gi in range(4):
        np.testing.assert_allclose(y[:, 0, 0, gi * 3:(gi + 1) * 3], x[:, 0, 0, gi * 8:(gi + 1) * 8] @ k[0, 0][:, gi * 3:(gi + 1) * 3],
                                   rtol=1e-12, atol=1e-14)


def test_shim_adamax_matches_torch_optim(tf):
    from tensorflow.keras.optimizers.experimental import Adamax
    rng = np.random.default_rng(1)
    w0 = rng.standard_normal(50)
    v = tf.Variable.make(w0.copy())
    opt = Adamax(learning_rate=2e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
    tw = torch.tensor(w0.copy(), requires_grad=True)
    topt = torch.optim.Adamax([tw], lr=2e-3, betas=(0.9, 0.999), eps=1e-8)
    for _ in range(5):
        gnp = rng.standard_normal(50)
        opt.apply_gradients([(tf.constant(gnp), v)])
        tw.grad = torch.tensor(gnp)
        topt.step()
    # torch puts eps inside the max, Keras adds it to the denominator: identical to ~eps/|g|
    np.testing.assert_allclose(v.numpy(), tw.detach().numpy(), rtol=0, atol=1e-8)
    assert np.abs(v.numpy() - w0).max() > 5e-3


def test_shim_clip_l2norm_bce_softmax(tf):
    rng = np.random.default_rng(2)
    g = rng.standard_normal((7, 3))
    n = np.linalg.norm(g)
    np.testing.assert_allclose(tf.clip_by_norm(tf.constant(g), 0.25).numpy(), g * 0.25 / n, rtol=1e-13)
    np.testing.assert_allclose(tf.clip_by_norm(tf.constant(g * 1e-3), 0.25).numpy(), g * 1e-3, rtol=1e-13)
    assert np.array_equal(tf.clip_by_norm(tf.constant(np.zeros(4)), 0.25).numpy(), np.zeros(4))
    np.testing.assert_allclose(tf.nn.l2_normalize(tf.constant(g), axis=None).numpy(), g / n, rtol=1e-13)
    x, z = rng.standard_normal(30) * 5, rng.uniform(0, 1, 30)
    want = torch.nn.functional.binary_cross_entropy_with_logits(torch.tensor(x), torch.tensor(z), reduction="none").numpy()
    np.testing.assert_allclose(tf.nn.sigmoid_cross_entropy_with_logits(labels=z, logits=x).numpy(), want, rtol=1e-12)
    e = np.exp(g - g.max(1, keepdims=True))
    np.testing.assert_allclose(tf.nn.softmax(tf.constant(g), axis=1).numpy(), e / e.sum(1, keepdims=True), rtol=1e-13)


def test_shim_weight_tracking_follows_keras_rules(tf):
    L = tf.keras.layers
    d = L.Dense(3)
    d.build((None, 5))
    assert [w.var_name for w in d.weights] == ["kernel", "bias"]
    d.kernel = None                                   # weight_norm.py:31: un-tracks the kernel variable
    assert [w.var_name for w in d.weights] == ["bias"]

    class Holder(L.Layer):
        def __init__(self):
            super().__init__()
            self.items = []
            self.own = self.add_weight("own", shape=[2])
            self.items.append(L.Dense(2))
            self.tail = L.Dense(1)
    h = Holder()
    h.items[0].build((None, 4)); h.tail.build((None, 4))
    assert [tuple(w.shape) for w in h.weights] == [(2,), (4, 2), (2,), (4, 1), (1,)]   # own first, then children in attribute order
    t = tf.constant(np.ones(3))
    t2 = t
    t2 += 1.0                                         # augmented assignment rebinds, never writes in place
    assert float(t.numpy().sum()) == 3.0 and float(t2.numpy().sum()) == 6.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree only exists in the build container")
def test_committed_fixture_is_what_the_reference_produces_now():
    root = os.path.dirname(HERE)
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "from oracle import make_golden_ref as m\n"
            "out = m.run_case('tiny_n9_m5', m._import_reference(), save=False)\n"
            "np.savez(sys.argv[1], **{k: v for k, v in out.items()})\n") % root
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "again.npz")
        subprocess.run([sys.executable, "-c", code, f], check=True, capture_output=True, cwd=root, timeout=300)
        again, g = np.load(f), np.load(os.path.join(HERE, "golden", "refexec_tiny_n9_m5.npz"))
        assert set(again.files) == set(g.files)
        for k in g.files:
            if g[k].dtype.kind in "fc":
                np.testing.assert_allclose(again[k], g[k], rtol=1e-12, atol=1e-14, err_msg=k)
            else:
                assert np.array_equal(again[k], g[k]), k


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree only exists in the build container")
def test_random_configurations_against_the_reference_executed_live():
    """Beyond the committed cases: random small configurations (widths, heads, K, nongt_dim, directions, label bias, residual,
    adaptive padding, init-like or trained-like weights) are run through the reference's own files in a subprocess and through
    the oracle here; logits, loss, q-mask, dq and every weight gradient must agree."""
    import json
    import subprocess
    import tempfile
    root = os.path.dirname(HERE)
    rng = np.random.default_rng(20261018)
    cases = {}
    for i in range(6):
        H = int(rng.choice([1, 2, 4]))
        kw = dict(v_dim=int(rng.choice([32, 48, 16 * H])), q_dim=int(rng.choice([16, 24])), rel_dim=16 * H, num_heads=H,
                  nongt_dim=int(rng.choice([3, 5, 8, 20])), num_answers=int(rng.integers(12, 40)), dir_num=int(rng.choice([1, 2])),
                  label_bias=bool(rng.integers(0, 2)), residual=bool(rng.integers(0, 2)))
        cases[f"rand{i}"] = (kw, int(rng.integers(1, 4)), int(rng.integers(2, 12)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), 1)
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r)\n"
            "from oracle import make_golden_ref as m\n"
            "cases = json.loads(sys.argv[1]); mods = m._import_reference()\n"
            "for name, c in cases.items():\n"
            "    m.CASES[name] = (c[0], c[1], c[2], c[3], c[4], c[5])\n"
            "    out = m.run_case(name, mods, save=False)\n"
            "    np.savez(sys.argv[2] + '/' + name + '.npz', **out)\n") % root
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([sys.executable, "-c", code, json.dumps(cases), tmp], capture_output=True, text=True, cwd=root, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        for name, (kw, B, N, adaptive, tl, _) in cases.items():
            g = np.load(os.path.join(tmp, name + ".npz"))
            cfg = HotPathConfig(**kw)
            inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=adaptive)
            named = syn.unflatten(cfg, syn.make_params(cfg, seed=7, trained_like=tl).astype(np.float64))
            loss, grads, dq_att, dq_last, out = ot.loss_and_grads(named, cfg, inp)
            np.testing.assert_allclose(out["logits"], g["logits"], rtol=1e-9, atol=1e-12, err_msg=name)
            np.testing.assert_allclose(loss, float(g["loss"]), rtol=1e-11, err_msg=name)
            np.testing.assert_allclose(dq_att, g["dq_att"], rtol=1e-8, atol=1e-13, err_msg=name)
            np.testing.assert_allclose(dq_last, g["dq_last"], rtol=1e-8, atol=1e-13, err_msg=name)
            f64 = lambda a: a.astype(np.float64)
            a = onp.forward(named, cfg, f64(inp["features"]), inp["boxes"], f64(inp["q_att"]), f64(inp["q_last"]), f64(inp["target"]))
            np.testing.assert_array_equal(a["mask"], g["mask"], err_msg=name)
            for i, e in enumerate(param_layout(cfg)[0]):
                np.testing.assert_allclose(grads[e.name], g["grad.full/" + e.name], rtol=0,
                                           atol=1e-9 * max(np.abs(grads[e.name]).max(), 1e-12) + 1e-14, err_msg=f"{name} {e.name}")
