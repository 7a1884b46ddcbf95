"""CPU-side checks of the drop-in boundary: libregat.so loads, exports every symbol include/regat.h declares, the ctypes
table covers them all, argument errors are reported through status codes + regat_last_error, and compute entry points
refuse to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tf_vqa_regat_b200 import _lib
from tf_vqa_regat_b200.config import HotPathConfig, param_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "regat.h")).read()
    return sorted(set(re.findall(r"REGAT_API\s+int\s+(regat_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    l = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/regat.h but not exported by libregat.so"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert l.regat_abi_version() == 1


def test_only_regat_symbols_are_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    syms = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert syms and all(s.startswith("regat_") for s in syms), [s for s in syms if not s.startswith("regat_")][:5]


def test_default_config_matches_reference_json():
    cfg = _lib.Config()
    assert _lib.lib().regat_default_config(C.byref(cfg)) == 0
    d = HotPathConfig()
    for f in ("v_dim", "q_dim", "rel_dim", "num_heads", "pos_emb_dim", "nongt_dim", "dir_num", "num_answers"):
        assert getattr(cfg, f) == getattr(d, f), f
    assert cfg.label_bias == 0 and cfg.residual == 1 and abs(cfg.grad_clip - 0.25) < 1e-7


@pytest.mark.parametrize("kw", [{}, dict(label_bias=True), dict(v_dim=1024), dict(dir_num=1)])
def test_engine_layout_equals_python_layout(kw):
    l = _lib.lib()
    cfg = HotPathConfig(**kw)
    from tf_vqa_regat_b200.engine import _c_config
    cc = _c_config(cfg)
    e = C.c_void_p()
    assert l.regat_engine_create(C.byref(cc), _lib.BF16, 8, 36, C.byref(e)) == 0
    pe, wb = C.c_int64(), C.c_int64()
    assert l.regat_engine_sizes(e, C.byref(pe), C.byref(wb)) == 0
    entries, total = param_layout(cfg)
    assert pe.value == total and wb.value > 0
    for i, en in enumerate(entries):
        off, n, la, k = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        assert l.regat_engine_param(e, i, C.byref(off), C.byref(n), C.byref(la), C.byref(k)) == 0
        assert (off.value, n.value, la.value, k.value) == (en.offset, en.numel, en.layer, {"v": 0, "g": 1, "b": 2}[en.kind])
    assert l.regat_engine_param(e, len(entries), None, None, None, None) == -1
    # unbound engine refuses to run
    assert l.regat_engine_forward(e, 1, 4, None, None, None, None, None, None, None) == -1
    assert "bind" in _lib.last_error()
    l.regat_engine_destroy(e)


def test_engine_create_rejects_unsupported_shapes():
    l = _lib.lib()
    from tf_vqa_regat_b200.engine import _c_config
    e = C.c_void_p()
    bad = [HotPathConfig(rel_dim=1000), HotPathConfig(pos_emb_dim=32), HotPathConfig(dir_num=3), HotPathConfig(v_dim=2050)]
    for cfg in bad:
        cc = _c_config(cfg)
        assert l.regat_engine_create(C.byref(cc), _lib.BF16, 8, 36, C.byref(e)) < 0
        assert _lib.last_error()
    cc = _c_config(HotPathConfig())
    assert l.regat_engine_create(C.byref(cc), 5, 8, 36, C.byref(e)) == -3
    assert l.regat_engine_create(C.byref(cc), _lib.F32, 8, 200, C.byref(e)) == -2


def test_argument_errors_without_gpu():
    l = _lib.lib()
    a = np.zeros(16, np.float32)
    assert l.regat_gemm(0, 0, 0, 2, 2, 2, None, 2, a.ctypes.data, 2, a.ctypes.data, 2, 0, None, None) == -1
    assert l.regat_position_embedding(None, 1, 2, 2, 64, None, None, None) == -1
    assert l.regat_position_embedding(a.ctypes.data, 1, 2, 2, 32, a.ctypes.data, a.ctypes.data, None) == -7
    assert l.regat_bce_fwd_bwd(2, 8, a.ctypes.data, 4, a.ctypes.data, a.ctypes.data, None, None, 8, 0, None) == -2
    msg = _lib.last_error()
    assert "leading dimension" in msg
    buf = C.create_string_buffer(8)
    n = l.regat_last_error(buf, 8)          # truncation is safe and NUL-terminated
    assert n == len(msg) and len(buf.value) == 7


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    l = _lib.lib()
    assert l.regat_device_count() == 0
    a = np.zeros(16, np.float32)
    assert l.regat_gemm(0, 0, 0, 2, 2, 2, a.ctypes.data, 2, a.ctypes.data, 2, a.ctypes.data, 2, 0, None, None) == -6
    assert "no CPU fallback" in _lib.last_error()
    from tf_vqa_regat_b200.engine import HotPathEngine
    with pytest.raises(_lib.RegatError):
        HotPathEngine(HotPathConfig(), 4, 36)
    from tf_vqa_regat_b200.model import prepare_graph_variables
    with pytest.raises(_lib.RegatError):
        prepare_graph_variables("implicit", np.zeros((1, 4, 4), np.float32), None, None, 4, 20, 64, 11, 15)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tf_vqa_regat_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_wave_divisors_follow_numpy_fp32():
    wd = _lib.wave_divisors(64)
    k = np.arange(0, 8, dtype=np.float32)
    np.testing.assert_array_equal(wd, np.power(np.full((1,), 1000, dtype=np.float32), (8.0 / 64) * k))
    assert wd.dtype == np.float32 and wd[0] == 1.0


def test_dp_exchange_argument_errors_need_no_gpu():
    """csrc/dp_exchange.cu validates before it launches: bad tables / ranks / alignment come back as status codes."""
    l = _lib.lib()
    ptrs = (C.c_uint64 * 2)(0x1000, 0x2000)
    assert l.regat_dp_allreduce_f32(None, 0, ptrs, 0, 2, 0, 1024, 1, 0, None) == -1
    assert l.regat_dp_allreduce_f32(ptrs, 0, ptrs, 2, 2, 0, 1024, 1, 0, None) == -1           # rank >= world
    assert l.regat_dp_allreduce_f32(ptrs, 0, ptrs, 0, 17, 0, 1024, 1, 0, None) == -1          # more ranks than the flag layout holds
    assert l.regat_dp_allreduce_f32(ptrs, 0, ptrs, 0, 2, 2, 1024, 1, 0, None) == -5           # offset not a multiple of 4 elements
    assert l.regat_dp_allreduce_f32(ptrs, 0, ptrs, 0, 2, 0, 0, 1, 0, None) == 0               # empty range: nothing to do
    assert l.regat_dp_reduce_bcast(ptrs, 0, ptrs, 0, 2, 4, 1024, 1, 0, None) == -5            # bf16 wire: multiples of 8
    assert l.regat_dp_wait_unpack(None, None, ptrs, 0, 2, 0, 1024, 1, None) == -1
    assert "dp_" in _lib.last_error()


def test_pad_ragged_argument_errors_need_no_gpu():
    l = _lib.lib()
    a = np.zeros(64, np.float32)
    off = np.zeros(3, np.int32)
    p, o = a.ctypes.data, off.ctypes.data
    assert l.regat_pad_ragged(2, 4, 6, 8, p, o, p, None, None) == -2            # width not a multiple of 4 floats
    assert "multiple of 4" in _lib.last_error()
    assert l.regat_pad_ragged(-1, 4, 4, 8, p, o, p, None, None) == -2
    assert l.regat_pad_ragged(0, 4, 4, 0, None, None, None, None, None) == 0     # empty batch: nothing to do
    assert l.regat_pad_ragged(2, 4, 4, 8, p, None, p, None, None) == -1          # null offsets
    assert l.regat_pad_ragged(2, 4, 4, 8, p + 4, o, p, None, None) == -5         # unaligned packed buffer
