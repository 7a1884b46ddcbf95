"""DLPack front door of the C ABI (regat_engine_forward_dl / regat_engine_train_step_dl, include/regat.h): what a TensorFlow
caller reaches through tf.experimental.dlpack.to_dlpack, exercised here with torch's capsules.
The library restates the DLPack v0.x structs; the first tests check that restatement against what torch really exports and
the argument validation, neither of which needs a GPU.  (Named zz so that it runs after every other GPU file under -x.)"""
import ctypes as C

import numpy as np
import pytest
import torch
from torch.utils.dlpack import to_dlpack

from tf_vqa_regat_b200 import _lib
from tf_vqa_regat_b200.config import HotPathConfig

SMALL = dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int32), ("dtype", DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class DLManagedTensor(C.Structure):
    _fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


_get = C.pythonapi.PyCapsule_GetPointer
_get.restype, _get.argtypes = C.c_void_p, [C.py_object, C.c_char_p]


def capsule_ptr(t):
    """(capsule, DLManagedTensor*) -- the capsule owns the export and must outlive the call (tensors are borrowed)."""
    cap = to_dlpack(t)
    return cap, _get(cap, b"dltensor")


def handmade(shape, device_type=2, device_id=0, code=2, bits=32, strides=None, data=0x1000, ndim=None):
    keep = {}
    keep["shape"] = (C.c_int64 * len(shape))(*shape)
    keep["strides"] = (C.c_int64 * len(shape))(*strides) if strides is not None else None
    m = DLManagedTensor()
    m.dl_tensor.data = data
    m.dl_tensor.device = DLDevice(device_type, device_id)
    m.dl_tensor.ndim = len(shape) if ndim is None else ndim
    m.dl_tensor.dtype = DLDataType(code, bits, 1)
    m.dl_tensor.shape = keep["shape"]
    m.dl_tensor.strides = keep["strides"] if strides is not None else C.POINTER(C.c_int64)()
    keep["m"] = m
    return keep, C.addressof(m)


def test_struct_restatement_matches_what_torch_exports():
    t = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)[:, 1:, :]          # non-trivial strides and offset
    cap, p = capsule_ptr(t)
    m = DLManagedTensor.from_address(p)
    d = m.dl_tensor
    assert d.device.device_type == 1 and d.ndim == 3                                      # kDLCPU
    assert (d.dtype.code, d.dtype.bits, d.dtype.lanes) == (2, 32, 1)                       # kDLFloat, 32 bits
    assert [d.shape[i] for i in range(3)] == [2, 2, 4] and [d.strides[i] for i in range(3)] == [12, 4, 1]
    assert d.data + d.byte_offset == t.data_ptr()
    del cap


def _engine():
    from tf_vqa_regat_b200.engine import _c_config
    l = _lib.lib()
    e = C.c_void_p()
    cc = _c_config(HotPathConfig(**SMALL))
    assert l.regat_engine_create(C.byref(cc), _lib.F32, 4, 36, C.byref(e)) == 0
    return l, e


def test_dl_argument_validation_needs_no_gpu():
    l, e = _engine()
    B, N = 2, 36
    ok = lambda shape, **k: handmade(shape, **k)
    kf, f = ok((B, N, 192)); kb, bx = ok((B, N, 4)); kq, qa = ok((B, 96)); kl, ql = ok((B, 96)); ko, lo = ok((B, 301))
    call = lambda *a: l.regat_engine_forward_dl(e, *a, None)
    assert call(None, bx, qa, ql, lo) == -1                                               # null tensor
    k1, cpu = ok((B, N, 192), device_type=1)
    assert call(cpu, bx, qa, ql, lo) == -4 and "CUDA device" in _lib.last_error()         # host tensor refused: no copy, no fallback
    k2, other = ok((B, N, 192), device_id=3)
    assert call(other, bx, qa, ql, lo) == -4
    k3, f64 = ok((B, N, 192), bits=64)
    assert call(f64, bx, qa, ql, lo) == -3
    k4, i32 = ok((B, N, 192), code=0)
    assert call(i32, bx, qa, ql, lo) == -3
    k5, flat = ok((B * N, 192))
    assert call(flat, bx, qa, ql, lo) == -2                                               # features must be [B,N,v_dim]
    k6, wrong_v = ok((B, N, 200))
    assert call(wrong_v, bx, qa, ql, lo) == -2
    k7, bad_boxes = ok((B, N, 6))
    assert call(f, bad_boxes, qa, ql, lo) == -2 and "boxes" in _lib.last_error()
    k8, bad_q = ok((B + 1, 96))
    assert call(f, bx, bad_q, ql, lo) == -2
    k9, strided = ok((B, N, 192), strides=(N * 192 * 2, 192, 1))
    assert call(strided, bx, qa, ql, lo) == -2 and "row-major" in _lib.last_error()
    k10, unaligned = ok((B, N, 192), data=0x1004)
    assert call(unaligned, bx, qa, ql, lo) == -5
    k11, bad_out = ok((B, 300))
    assert call(f, bx, qa, ql, bad_out) == -2 and "logits_out" in _lib.last_error()
    k12, compact = ok((B, N, 192), strides=(N * 192, 192, 1))                             # explicit compact strides are fine
    assert call(compact, bx, qa, ql, lo) == -1 and "bind" in _lib.last_error()            # everything valid: reaches the (unbound) engine
    loss = np.zeros(2, np.float32)
    k13, tg = ok((B, 301))
    assert l.regat_engine_train_step_dl(e, f, bx, qa, ql, tg, 1e-3, 1, loss.ctypes.data, None) == -1
    k14, bad_t = ok((B, 302))
    assert l.regat_engine_train_step_dl(e, f, bx, qa, ql, bad_t, 1e-3, 1, loss.ctypes.data, None) == -2
    l.regat_engine_destroy(e)


def test_torch_cpu_capsules_are_refused():
    l, e = _engine()
    ts = [torch.zeros(2, 36, 192), torch.zeros(2, 36, 4), torch.zeros(2, 96), torch.zeros(2, 96), torch.zeros(2, 301)]
    caps = [capsule_ptr(t) for t in ts]
    assert l.regat_engine_forward_dl(e, *[p for _, p in caps], None) == -4
    l.regat_engine_destroy(e)


@pytest.mark.gpu
def test_dl_front_door_equals_pointer_front_door():
    """Zero copy: the DLPack entry points read the very buffers torch exported and give the same bits as the raw-pointer ones."""
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.engine import HotPathEngine
    cfg = HotPathConfig(**SMALL)
    B, N = 3, 36
    eng = HotPathEngine(cfg, B, N, dtype="fp32")
    eng.load_params(syn.make_params(cfg, seed=7, trained_like=True))
    inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=True)
    d = {k: torch.tensor(v).cuda() for k, v in inp.items() if k != "n_obj"}
    want = eng.forward(d["features"], d["boxes"], d["q_att"], d["q_last"])
    out = torch.empty(B, cfg.num_answers, device="cuda")
    caps = [capsule_ptr(t) for t in (d["features"], d["boxes"], d["q_att"], d["q_last"], out)]
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(eng.lib.regat_engine_forward_dl(eng._h, *[p for _, p in caps], st))
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    # a training step through DLPack moves the parameters exactly like the pointer call does on a twin engine
    twin = HotPathEngine(cfg, B, N, dtype="fp32")
    twin.load_params(syn.make_params(cfg, seed=7, trained_like=True))
    l1 = twin.train_step(d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"], 1e-3, 1).clone()
    loss = torch.zeros(2, device="cuda")
    caps = [capsule_ptr(t) for t in (d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"])]
    _lib.check(eng.lib.regat_engine_train_step_dl(eng._h, *[p for _, p in caps], 1e-3, 1, loss.data_ptr(), st))
    torch.cuda.synchronize()
    assert abs(float(loss[0]) - float(l1[0])) <= 1e-5 * abs(float(l1[0]))
    # same update on both engines.  Not compared bit for bit: reductions that use atomics may order their sums differently from
    # run to run, and where a gradient element is pure rounding noise Adamax turns its sign into a +-lr step
    assert (eng.params - twin.params).abs().mean().item() < 1e-6
    assert (eng.params - torch.tensor(syn.make_params(cfg, seed=7, trained_like=True)).cuda()).abs().max().item() > 5e-4
    # a tensor on the wrong device type is refused with a status, not copied
    host_features = torch.tensor(inp["features"])
    cpu_cap, out_cap = capsule_ptr(host_features), capsule_ptr(out)
    assert eng.lib.regat_engine_forward_dl(eng._h, cpu_cap[1], caps[1][1], caps[2][1], caps[3][1], out_cap[1], st) == -4
