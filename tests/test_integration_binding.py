"""integration/regat_b200_tf.py -- the binding a maintainer of the reference adds -- executed against a
stand-in `tensorflow` module (tests/fake_tf: torch-backed tensors, real DLPack capsules).  CPU: the module imports, parses
DLPack capsules correctly and refuses to run without a device.  GPU: logits and train steps through the binding equal the
repository's own engine wrapper on the same weights and inputs."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture()
def binding(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(HERE, "fake_tf"))
    monkeypatch.syspath_prepend(os.path.join(ROOT, "integration"))
    for m in [m for m in sys.modules if m == "tensorflow" or m.startswith("tensorflow.")] + ["regat_b200_tf"]:
        sys.modules.pop(m, None)
    mod = importlib.import_module("regat_b200_tf")
    yield mod
    for m in [m for m in sys.modules if m == "tensorflow" or m.startswith("tensorflow.")] + ["regat_b200_tf"]:
        sys.modules.pop(m, None)


def test_binding_has_no_elided_bodies():
    src = open(os.path.join(ROOT, "integration", "regat_b200_tf.py")).read()
    assert "..." not in src.replace("[...]", ""), "the binding must be complete: no ellipsis bodies"
    for name in ("regat_engine_forward_dl", "regat_engine_train_step_dl", "regat_engine_bind", "regat_engine_params_changed"):
        assert name in src


def test_dlpack_struct_layout_matches_a_real_capsule(binding):
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)[:, 1:, :]          # non-trivial view: offset into the storage
    t = t.contiguous()
    b = binding.Borrowed(t)
    assert b.shape == (2, 2, 4)
    assert b.data == t.data_ptr()
    assert not b.on_gpu
    d = b.managed.dl_tensor
    assert (d.dtype.code, d.dtype.bits, d.dtype.lanes) == (2, 32, 1)              # kDLFloat, 32 bits
    assert d.ndim == 3 and d.device.device_type == 1                               # kDLCPU


def test_config_fields_match_the_header(binding):
    import re
    hdr = open(os.path.join(ROOT, "include", "regat.h")).read()
    body = hdr[hdr.index("typedef struct regat_config"):hdr.index("} regat_config;")]
    names = re.findall(r"\b(?:int32_t|float)\s+([a-z_0-9, ]+);", body)
    fields = [x.strip() for n in names for x in n.split(",")]
    assert fields == [f for f, _ in binding.RegatConfig._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a device")
def test_refuses_to_run_without_a_device(binding):
    with pytest.raises(RuntimeError, match="no CUDA device"):
        binding.ReGATEngine(dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, num_answers=301), max_batch=2, max_rois=36, dtype="fp32")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_binding_equals_engine_wrapper(binding, dtype):
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.config import HotPathConfig, param_layout
    from tf_vqa_regat_b200.engine import HotPathEngine
    kw = dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)
    cfg = HotPathConfig(**kw)
    B, N, lr = 4, 36, 1e-3
    flat = syn.make_params(cfg, seed=7, trained_like=True)
    entries, _ = param_layout(cfg)
    arrays = [flat[e.offset:e.offset + e.numel].reshape(e.shape) for e in entries]
    inp = [syn.make_inputs(cfg, B, N, seed=1000 + s, adaptive=True) for s in range(2)]
    dev = [{k: torch.tensor(v).cuda() for k, v in b.items() if k != "n_obj"} for b in inp]
    ref = HotPathEngine(cfg, B, N, dtype=dtype)
    ref.load_params(flat)
    eng = binding.ReGATEngine(dict(kw, label_bias=int(cfg.label_bias), residual=int(cfg.residual)), max_batch=B, max_rois=N, dtype=dtype)
    assert [(o, n) for o, n, _, _ in eng.layout] == [(e.offset, e.numel) for e in entries]
    eng.set_weights(arrays)
    d = dev[0]
    a = eng.logits(d["features"], d["boxes"], d["q_att"], d["q_last"])
    b = ref.forward(d["features"], d["boxes"], d["q_att"], d["q_last"])
    assert torch.equal(a, b)                                       # same kernels, same weights, same inputs
    for s in range(3):
        d = dev[s & 1]
        loss, score = eng.train_step(d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"], lr)
        want = ref.train_step(d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"], lr, s + 1)
        assert abs(loss - float(want[0])) < (2e-5 if dtype == "fp32" else 2e-3) * abs(float(want[0])), (s, loss, float(want[0]))
    got = np.concatenate([w.ravel() for w in eng.get_weights()])
    want = np.concatenate([ref.params[e.offset:e.offset + e.numel].cpu().numpy().ravel() for e in entries])
    dd = np.abs(got - want)
    assert dd.max() <= 2 * lr + 1e-6 and dd.mean() < 0.02 * lr      # atomics reorder additions; Adamax turns sign noise into +-lr
    # shape errors surface as Python exceptions, not as launches
    with pytest.raises(ValueError):
        eng.logits(d["features"][:, :, :8].contiguous(), d["boxes"], d["q_att"], d["q_last"])
    with pytest.raises(ValueError):
        eng.set_weights(arrays[:-1])

    class _Var:                                                     # a Keras variable as far as load_from_keras is concerned
        def __init__(self, a): self.a = a; self.shape = a.shape
        def numpy(self): return self.a

    class _Sub:
        def __init__(self, arrs): self.weights = [_Var(a) for a in arrs]
        def set_weights(self, arrs): self.weights = [_Var(np.asarray(a)) for a in arrs]

    names = [e.name for e in entries]
    split = [i for i, n in enumerate(names) if n.startswith("joint_emb.")][0], [i for i, n in enumerate(names) if n.startswith("classifier.")][0]

    class _Model:
        v_relation = _Sub(arrays[:split[0]]); joint_emb = _Sub(arrays[split[0]:split[1]]); classifier = _Sub(arrays[split[1]:])

    eng.load_from_keras(_Model)
    back = eng.get_weights()
    assert all(np.array_equal(x.ravel(), y.ravel()) for x, y in zip(back, arrays))
    eng.store_to_keras(_Model)
    assert np.array_equal(_Model.classifier.weights[-1].numpy(), arrays[-1])
    eng.close()
