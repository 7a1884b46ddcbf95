"""The whole train step as ONE CUDA graph (regat_engine_train_step_dev: device-resident lr / step counter, per-range clip +
Adamax and re-derived bf16 kernels on the engine's optimizer stream) against the eager host-scalar step.  A captured step
must (a) really consume the parameters the previous replay wrote and (b) follow parameters written from outside."""
import numpy as np
import pytest
import torch

from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig, param_layout

pytestmark = pytest.mark.gpu
SMALL = dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)


def _zero_direction(name):
    return ("implicit_relation.bias/" in name or name.endswith(".key/bias")
            or name in ("joint_emb.linear/bias", "joint_emb.v2attention/bias"))


def _engines(dtype, B=4, N=36, kw=SMALL):
    from tf_vqa_regat_b200.engine import HotPathEngine
    cfg = HotPathConfig(**kw)
    flat = syn.make_params(cfg, seed=7, trained_like=True)
    batches = [{k: torch.tensor(v).cuda() for k, v in syn.make_inputs(cfg, B, N, seed=1000 + s, adaptive=True).items() if k != "n_obj"}
               for s in range(2)]
    a, b = HotPathEngine(cfg, B, N, dtype=dtype), HotPathEngine(cfg, B, N, dtype=dtype)
    a.load_params(flat); b.load_params(flat)
    return cfg, flat, batches, a, b


def _args(d):
    return d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_graph_replay_equals_eager_steps(dtype):
    cfg, flat, batches, eager, graphed = _engines(dtype)
    lr, steps = 2e-3, 6
    losses_e = []
    for s in range(steps):
        losses_e.append(float(eager.train_step(*_args(batches[s & 1]), lr, s + 1)[0]))
    graphs = [graphed.capture_train_step(*_args(batches[k])) for k in range(2)]
    graphed.set_lr(lr)
    graphed.set_step(0)
    losses_g = []
    for s in range(steps):
        losses_g.append(float(graphs[s & 1].replay()[0]))
    torch.cuda.synchronize()
    assert graphed.get_step() == (steps, pytest.approx(lr))
    # every replay saw the parameters the previous one wrote: same losses as the eager loop, and they move
    np.testing.assert_allclose(losses_g, losses_e, rtol=2e-5 if dtype == "fp32" else 2e-3)
    assert abs(losses_e[2] - losses_e[0]) > 1e-3 * abs(losses_e[0]), "the parameters never changed"
    pe, pg = eager.params.cpu().numpy(), graphed.params.cpu().numpy()
    for e in param_layout(cfg)[0]:
        if _zero_direction(e.name):
            continue
        d = np.abs(pe[e.offset:e.offset + e.numel] - pg[e.offset:e.offset + e.numel])
        # same kernels in the same order; only atomics (split-K, bias sums) reorder additions.  Adamax turns a gradient whose
        # sign that reordering decides (g ~ 0) into a +-lr move, so single elements may differ by a step; on average nothing does
        assert d.max() <= 2.0 * lr + 1e-6 and (d.mean() < 0.02 * lr or e.numel == 1), (e.name, d.max(), d.mean())
    # Adamax slots follow too (the trajectories drift apart by the sign decisions above, hence a bound relative to the largest slot)
    du = float((eager.adamax_u - graphed.adamax_u).abs().max()) / float(eager.adamax_u.abs().max())
    assert du < (1e-4 if dtype == "fp32" else 2e-2), du


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_captured_step_follows_outside_parameter_writes(dtype):
    """load_params after the capture: the replay must train the NEW parameters (the derived bf16 kernels / alpha are rebuilt when
    the parameters are written, not by a host-side shortcut baked into the graph)."""
    cfg, flat, batches, eager, graphed = _engines(dtype)
    lr = 1e-3
    g = graphed.capture_train_step(*_args(batches[0]))
    graphed.set_lr(lr); graphed.set_step(0)
    g.replay(); g.replay()
    other = syn.make_params(cfg, seed=11, trained_like=True)
    graphed.load_params(other)
    graphed.adamax_m.zero_(); graphed.adamax_u.zero_(); graphed.set_step(0)
    eager.load_params(other)
    le = float(eager.train_step(*_args(batches[0]), lr, 1)[0])
    lg = float(g.replay()[0])
    assert abs(le - lg) < (2e-5 if dtype == "fp32" else 2e-3) * abs(le)
    le2 = float(eager.train_step(*_args(batches[0]), lr, 2)[0])
    lg2 = float(g.replay()[0])
    assert abs(le2 - lg2) < (2e-5 if dtype == "fp32" else 2e-3) * abs(le2)
    assert abs(le2 - le) > 1e-4 * abs(le)


def test_update_keeps_derived_state_current_bf16():
    """After regat_engine_update the bf16 kernels / alpha are those of the NEW parameters without any per-forward prep: a
    forward-only call right after an update equals a fresh engine loaded with the updated parameters."""
    from tf_vqa_regat_b200.engine import HotPathEngine
    cfg, flat, batches, eng, _ = _engines("bf16")
    d = batches[0]
    eng.fwd_bwd(*_args(d))
    eng.update(1e-2, 1)
    n0 = None
    after = eng.forward(d["features"], d["boxes"], d["q_att"], d["q_last"])
    n0 = eng.last_launches()
    again = eng.forward(d["features"], d["boxes"], d["q_att"], d["q_last"])
    assert eng.last_launches() == n0, "a forward pass right after an update must not need any weight preparation"
    assert torch.equal(after, again)
    fresh = HotPathEngine(cfg, 4, 36, dtype="bf16")
    fresh.load_params(eng.params.clone())
    want = fresh.forward(d["features"], d["boxes"], d["q_att"], d["q_last"])
    rel = float((after - want).abs().max() / want.abs().max())
    assert rel < 1e-5, rel
