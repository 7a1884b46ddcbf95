"""The CUDA path (through the C ABI) against vectors produced by executing the reference's own model/*.py and
train.py (tests/golden/refexec_*.npz, oracle/make_golden_ref.py).  No oracle in between: fixture in, kernels, compare.
Tolerances are north_star's: q-mask exact, fp32 mode 1e-4 relative, bf16 mode 1e-2 on logits with the same answer."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig, param_layout

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
# the kernels need head dim 64: the small_* and full_* fixtures qualify, tiny_* do not
ORDER = ["small_n36_m20", "small_n12_clamped", "small_dir1_labelbias_nores", "small_n36_m36_fullkk", "small_n100_m20_adaptive",
         "small_nov2out", "small_n36_m20_init", "full_b4_n36_m20"]
FILES = [os.path.join(HERE, "golden", f"refexec_{n}.npz") for n in ORDER]
IDS = ORDER


def test_every_kernel_sized_fixture_is_covered():
    have = {os.path.basename(f)[len("refexec_"):-4] for f in glob.glob(os.path.join(HERE, "golden", "refexec_*.npz"))}
    assert {n for n in have if not n.startswith(("tiny_", "question_", "collate"))} == set(ORDER)   # question_*: tests/test_refexec_question.py


def _zero_direction(name):
    return ("implicit_relation.bias/" in name or name.endswith(".key/bias")
            or name in ("joint_emb.linear/bias", "joint_emb.v2attention/bias"))


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _load(path, dtype):
    from tf_vqa_regat_b200.engine import HotPathEngine
    g = np.load(path)
    cfg = HotPathConfig(**ast.literal_eval(str(g["cfg"])))
    B, N, steps = int(g["B"]), int(g["N"]), int(g["steps"])
    batches = [syn.make_inputs(cfg, B, N, seed=1000 + s, adaptive=bool(g["adaptive"])) for s in range(steps + 1)]
    np.testing.assert_allclose([float(np.sum(b["features"], dtype=np.float64)) for b in batches], g["input_check"], rtol=1e-12)
    eng = HotPathEngine(cfg, B, N, dtype=dtype)
    eng.load_params(syn.make_params(cfg, seed=7, trained_like=bool(g["trained_like"])))
    dev = [{k: torch.tensor(v).cuda() for k, v in b.items() if k != "n_obj"} for b in batches]
    return g, cfg, B, N, steps, eng, dev, batches


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_forward_fp32_vs_reference_execution(path):
    g, cfg, B, N, steps, eng, dev, _ = _load(path, "fp32")
    d = dev[0]
    logits, att = eng.forward(d["features"], d["boxes"], d["q_att"], d["q_last"], return_att=True)
    np.testing.assert_array_equal(eng.buffer("mask", (B, N), torch.float32).cpu().numpy(), g["mask"])   # bit-exact
    v1 = eng.buffer("v1", (B, N, cfg.rel_dim)).cpu().numpy()
    if "v1" in g:
        assert _rel(v1, g["v1"]) < 1e-4
    else:
        assert np.abs(v1[:, :, :16] - g["v1_head"]).max() < 1e-4 * np.abs(v1).max()
    assert _rel(att.cpu().numpy(), g["att_weights"]) < 1e-4
    assert _rel(logits.cpu().numpy(), g["logits"]) < 1e-4
    assert np.array_equal(logits.argmax(1).cpu().numpy(), g["logits"].argmax(1))


@pytest.mark.parametrize("path", [f for f in FILES if "nov2out" not in f and "dir1" not in f],
                         ids=[i for i in IDS if "nov2out" not in i and "dir1" not in i])
def test_forward_bf16_vs_reference_execution(path):
    g, cfg, B, N, steps, eng, dev, _ = _load(path, "bf16")
    d = dev[0]
    logits = eng.forward(d["features"], d["boxes"], d["q_att"], d["q_last"]).cpu().numpy()
    ref = g["logits"]
    assert _rel(logits, ref) < 1e-2
    top2 = np.sort(ref, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 2e-2 * np.abs(ref).max()          # same answer wherever bf16 can resolve it
    assert np.array_equal(logits.argmax(1)[clear], ref.argmax(1)[clear])


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_gradients_fp32_vs_reference_tape(path):
    g, cfg, B, N, steps, eng, dev, batches = _load(path, "fp32")
    d = dev[0]
    out = eng.fwd_bwd(d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"], want_dq=True, want_logits=True)
    eng.finalize_grads()
    torch.cuda.synchronize()
    loss = float(g["loss"])
    assert abs(float(out["loss"]) - loss) < 1e-4 * abs(loss)
    assert _rel(out["dq_att"].cpu().numpy(), g["dq_att"]) < 2e-4
    assert _rel(out["dq_last"].cpu().numpy(), g["dq_last"]) < 2e-4
    got = {k: v.cpu().numpy().astype(np.float64) for k, v in eng.named(eng.grads).items()}
    scale = max(float(g["grad.absmax/joint_emb.linear/v"]), 1e-12)
    bad = {}
    for i, e in enumerate(param_layout(cfg)[0]):
        a, name = got[e.name], e.name
        if _zero_direction(name):
            assert np.abs(a).max() < 1e-3 * scale + 1e-6, (name, np.abs(a).max())      # rounding noise on both sides
            continue
        # pair_pos_fc gradients carry dL/z with z ~ 0 entries ("geometry noise")
        tol = 2e-2 if "pair_pos_fc" in name else 5e-4
        absmax, norm = max(float(g[f"grad.absmax/{name}"]), 1e-30), float(g[f"grad.norm/{name}"])
        err = np.abs(a.ravel()[g[f"grad.idx/{name}"]] - g[f"grad.sample/{name}"]).max() / absmax
        nerr = abs(np.sqrt((a * a).sum()) - norm) / max(norm, 1e-30)
        r = np.random.default_rng(100 + i).standard_normal(a.size)
        perr = abs(a.ravel() @ r - float(g[f"grad.proj/{name}"])) / max(norm, 1e-30)
        if err > tol or nerr > 10 * tol or perr > 50 * tol:
            bad[name] = (err, nerr, perr)
    assert not bad, bad


@pytest.mark.parametrize("path", [f for f in FILES if "full_b4" not in f], ids=[i for i in IDS if "full_b4" not in i])
def test_train_steps_fp32_vs_reference_train_loop(path):
    """The reference's train.train() ran `steps` batches (GradientTape, per-tensor clip_by_norm, Adamax, lr 1e-3) and then
    train.evaluate() on one more: per-step losses, parameters afterwards, evaluation logits."""
    g, cfg, B, N, steps, eng, dev, batches = _load(path, "fp32")
    lr = float(g["lr"])
    for s in range(steps):
        d = dev[s]
        l = eng.train_step(d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"], lr, s + 1)
        x, z = g["train.logits"][s], batches[s]["target"].astype(np.float64)
        want = (np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))).mean() * z.shape[1]      # train.py:23,107-108
        assert abs(float(l[0]) - want) < 2e-4 * abs(want), (s, float(l[0]), want)
    got = {k: v.cpu().numpy().astype(np.float64) for k, v in eng.named().items()}
    for e in param_layout(cfg)[0]:
        if _zero_direction(e.name):
            continue     # Adamax normalises pure rounding noise to +-lr there
        a = got[e.name].ravel()[g[f"param.idx/{e.name}"]]
        # an Adamax step is at most lr per element: agreement to a small fraction of the total movement
        assert np.abs(a - g[f"param.sample/{e.name}"]).max() < 0.05 * steps * lr + 1e-6, e.name
    d = dev[steps]
    logits = eng.forward(d["features"], d["boxes"], d["q_att"], d["q_last"]).cpu().numpy()
    assert _rel(logits, g["eval.logits"]) < 1e-2
