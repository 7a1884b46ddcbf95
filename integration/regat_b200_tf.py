"""model/regat_b200.py -- the file a maintainer of jhss/TF_VQA_ReGAT adds to run the implicit-relation hot path on libregat.so.

It replaces, for relation_type == 'implicit' and fusion == 'butd':
    rel_graph_net.py:53-62   v_relation -> joint_emb -> classifier          -> ReGATEngine.logits
    train.py:97              prepare_graph_variables (host NumPy, 47-655 MB) -> not needed: the kernels take the boxes
    train.py:103-113         GradientTape, per-tensor clip_by_norm, Adamax   -> ReGATEngine.train_step
and leaves the language front-end, the dataset, logging and the CLI of the reference untouched (INTEGRATION.md shows the
call-site diff).  Tensors cross the boundary as DLPack capsules (tf.experimental.dlpack.to_dlpack), zero copy; the library
allocates nothing -- parameters, gradients, Adamax slots and the workspace are TensorFlow tensors created here and bound once.

TensorFlow eager exposes no CUDA stream, so every call is bracketed by a device synchronisation and runs on the NULL stream.
Only `tensorflow`, `numpy` and `ctypes` are needed.  (The repository's own tests run this file against a stand-in
`tensorflow` module that hands out torch-backed DLPack capsules, tests/fake_tf, because TensorFlow is not installable there.)
"""
import ctypes as C
import os

import numpy as np
import tensorflow as tf
from tensorflow.experimental import dlpack as tfdl

F32, BF16 = 0, 1
_CFG_INT = ("v_dim", "q_dim", "rel_dim", "num_heads", "pos_emb_dim", "nongt_dim", "dir_num", "num_answers", "label_bias", "residual")
_CFG_FLT = ("grad_clip", "beta1", "beta2", "eps")


class RegatConfig(C.Structure):                       # include/regat.h: regat_config
    _fields_ = [(n, C.c_int32) for n in _CFG_INT] + [(n, C.c_float) for n in _CFG_FLT]


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):                         # dlpack.h: DLTensor
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):                  # dlpack.h: DLManagedTensor
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


_PyCapsule_GetPointer = C.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype, _PyCapsule_GetPointer.argtypes = C.c_void_p, [C.py_object, C.c_char_p]


class Borrowed:
    """A tf.Tensor lent to the library for one call: the capsule keeps the buffer alive, `ptr` is its DLManagedTensor*."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.capsule = tfdl.to_dlpack(tensor)
        self.ptr = _PyCapsule_GetPointer(self.capsule, b"dltensor")
        self.managed = C.cast(self.ptr, C.POINTER(_DLManagedTensor)).contents

    @property
    def data(self):
        t = self.managed.dl_tensor
        return (t.data or 0) + t.byte_offset

    @property
    def shape(self):
        t = self.managed.dl_tensor
        return tuple(int(t.shape[i]) for i in range(t.ndim))

    @property
    def on_gpu(self):
        return self.managed.dl_tensor.device.device_type == 2          # kDLCUDA


def _find_library(path=None):
    for cand in (path, os.environ.get("REGAT_LIB"),
                 os.path.join(os.path.dirname(os.path.abspath(__file__)), "libregat.so"),
                 os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tf_vqa_regat_b200", "libregat.so")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("libregat.so not found: build it (python -m tf_vqa_regat_b200.build) and pass its path or set REGAT_LIB; "
                       "there is no CPU or TensorFlow fallback for this path")


class ReGATEngine:
    """The hot path of one GPU behind the reference's tensors.

    cfg: the reference's hyper-parameters (config/butd_vqa.json names -> regat_config fields), e.g.
         dict(v_dim=2048, q_dim=768, rel_dim=1024, num_heads=16, nongt_dim=20, dir_num=2, num_answers=3129, label_bias=0)
    dtype: "bf16" (tcgen05 GEMMs, fp32 master weights; 1e-2 on logits) or "fp32" (parity kernels, 1e-4)."""

    def __init__(self, cfg=None, max_batch=256, max_rois=100, dtype="bf16", lib_path=None):
        self.lib = C.CDLL(_find_library(lib_path))
        L = self.lib
        for name in ("regat_engine_create", "regat_engine_destroy", "regat_engine_sizes", "regat_engine_param", "regat_engine_bind",
                     "regat_engine_forward_dl", "regat_engine_train_step_dl", "regat_engine_params_changed", "regat_default_config",
                     "regat_last_error", "regat_memcpy", "regat_device_synchronize", "regat_abi_version", "regat_device_count"):
            getattr(L, name).restype = C.c_int
        L.regat_engine_create.argtypes = [C.POINTER(RegatConfig), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.regat_engine_destroy.argtypes = [C.c_void_p]
        L.regat_engine_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.regat_engine_param.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.regat_engine_bind.argtypes = [C.c_void_p] * 6 + [C.c_int64]
        L.regat_engine_forward_dl.argtypes = [C.c_void_p] * 7
        L.regat_engine_train_step_dl.argtypes = [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.regat_engine_params_changed.argtypes = [C.c_void_p]
        L.regat_default_config.argtypes = [C.POINTER(RegatConfig)]
        L.regat_last_error.argtypes = [C.c_char_p, C.c_size_t]
        L.regat_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        if L.regat_abi_version() != 1:
            raise RuntimeError("libregat.so: ABI version mismatch")
        self.h = C.c_void_p()
        if L.regat_device_count() == 0:
            raise RuntimeError("regat: no CUDA device -- the hot path has no CPU fallback")
        self.cfg = RegatConfig()
        self._check(L.regat_default_config(C.byref(self.cfg)))
        for k, v in (cfg or {}).items():
            if k not in _CFG_INT + _CFG_FLT:
                raise ValueError(f"unknown configuration field '{k}'")
            setattr(self.cfg, k, type(getattr(self.cfg, k))(v))
        self.dtype = {"bf16": BF16, "fp32": F32}[dtype]
        self.max_batch, self.max_rois = int(max_batch), int(max_rois)
        self._check(L.regat_engine_create(C.byref(self.cfg), self.dtype, self.max_batch, self.max_rois, C.byref(self.h)))
        n, ws = C.c_int64(), C.c_int64()
        self._check(L.regat_engine_sizes(self.h, C.byref(n), C.byref(ws)))
        self.param_elems, self.workspace_bytes = n.value, ws.value
        # flat fp32 buffers (Keras variable order, 64-element aligned entries) and the workspace: TensorFlow tensors, lent for the
        # lifetime of the engine (the capsules are kept so that TensorFlow cannot recycle the memory)
        with tf.device("/GPU:0"):
            self._bufs = [Borrowed(tf.zeros([self.param_elems], tf.float32)) for _ in range(4)]
            self._ws = Borrowed(tf.zeros([(self.workspace_bytes + 3) // 4 + 64], tf.float32))
            self._loss = Borrowed(tf.zeros([2], tf.float32))
        self._sync()
        self.params, self.grads, self.adamax_m, self.adamax_u = self._bufs
        wsp = (self._ws.data + 255) & ~255
        if any(b.data & 255 for b in self._bufs):
            raise RuntimeError("regat: TensorFlow returned a parameter buffer that is not 256-byte aligned")
        self._check(L.regat_engine_bind(self.h, self.params.data, self.grads.data, self.adamax_m.data, self.adamax_u.data, wsp,
                                        self.workspace_bytes))
        self.layout = []                                   # (offset, numel, layer index, kind 0=v 1=g 2=bias) in Keras variable order
        i = 0
        off, num, lay, kind = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        while L.regat_engine_param(self.h, i, C.byref(off), C.byref(num), C.byref(lay), C.byref(kind)) == 0:
            self.layout.append((off.value, num.value, lay.value, kind.value))
            i += 1
        self.steps_taken = 0

    # ---- plumbing
    def _check(self, status):
        if status != 0:
            buf = C.create_string_buffer(512)
            self.lib.regat_last_error(buf, 512)
            raise RuntimeError(f"libregat status {status}: {buf.value.decode(errors='replace')}")

    def _sync(self):
        self._check(self.lib.regat_device_synchronize())

    def close(self):
        if getattr(self, "h", None):
            self._sync()
            self.lib.regat_engine_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- weights: Keras by-order lists, exactly what model.get_weights() / set_weights() of the three sub-models use
    def set_weights(self, arrays):
        """arrays: the variables of v_relation, joint_emb and classifier in Keras order (per WeightNorm: v, g, bias)."""
        arrays = list(arrays)
        if len(arrays) != len(self.layout):
            raise ValueError(f"set_weights: expected {len(self.layout)} arrays, got {len(arrays)}")
        flat = np.zeros(self.param_elems, dtype=np.float32)
        for (off, num, _, _), a in zip(self.layout, arrays):
            a = np.asarray(a, dtype=np.float32)
            if a.size != num:
                raise ValueError(f"set_weights: variable at offset {off} has {num} elements, got an array of shape {a.shape}")
            flat[off:off + num] = a.ravel()
        self._sync()
        self._check(self.lib.regat_memcpy(self.params.data, flat.ctypes.data, flat.nbytes, 1, None))
        self._check(self.lib.regat_engine_params_changed(self.h))

    def get_weights(self, shapes=None):
        """Inverse of set_weights; `shapes` (optional, by order) reshapes the flat slices."""
        flat = np.empty(self.param_elems, dtype=np.float32)
        self._sync()
        self._check(self.lib.regat_memcpy(flat.ctypes.data, self.params.data, flat.nbytes, 2, None))
        out = [flat[off:off + num].copy() for off, num, _, _ in self.layout]
        if shapes is not None:
            out = [a.reshape(s) for a, s in zip(out, shapes)]
        return out

    def load_from_keras(self, model):
        """model: the reference's RelationGraphAttentionNetwork after its variables exist (rel_graph_net.py:113-123)."""
        arrays = []
        for sub in (model.v_relation, model.joint_emb, model.classifier):
            arrays.extend(v.numpy() for v in sub.weights)
        self.set_weights(arrays)

    def store_to_keras(self, model):
        """Writes the trained parameters back into the Keras sub-models (so that model.save_weights, main.py:145, keeps working)."""
        subs = (model.v_relation, model.joint_emb, model.classifier)
        shapes = [tuple(v.shape) for sub in subs for v in sub.weights]
        arrays = self.get_weights(shapes)
        i = 0
        for sub in subs:
            n = len(sub.weights)
            sub.set_weights(arrays[i:i + n])
            i += n

    # ---- the two calls
    def _lend(self, t, shape, what):
        if t.dtype != tf.float32:
            raise TypeError(f"{what}: expected float32, got {t.dtype}")
        b = Borrowed(t)
        if not b.on_gpu:
            raise RuntimeError(f"{what}: tensor is not in GPU memory (place the batch with tf.device('/GPU:0'))")
        if b.shape != tuple(shape):
            raise ValueError(f"{what}: expected shape {tuple(shape)}, got {b.shape}")
        return b

    def _inputs(self, visual, bb, q_att, q_last):
        B, N = int(visual.shape[0]), int(visual.shape[1])
        if B > self.max_batch or N > self.max_rois:
            raise ValueError(f"batch {B}x{N} exceeds the engine capacity {self.max_batch}x{self.max_rois}")
        c = self.cfg
        return B, N, [self._lend(visual, (B, N, c.v_dim), "visual"), self._lend(bb, (B, N, 4), "bounding boxes"),
                      self._lend(q_att, (B, c.q_dim), "q_emb_self_att"), self._lend(q_last, (B, c.q_dim), "q_emb")]

    def logits(self, visual, bb, q_att, q_last):
        """rel_graph_net.py:53-62: visual [B,N,v_dim], bb [B,N,4] absolute pixels (x1,y1,x2,y2) -- the array train.py:94-97 feeds to
        prepare_graph_variables -- q_att = q_emb_self_att, q_last = q_emb.  Returns logits [B, num_answers]."""
        B, N, lent = self._inputs(visual, bb, q_att, q_last)
        with tf.device("/GPU:0"):
            out = tf.zeros([B, self.cfg.num_answers], tf.float32)
        lent.append(self._lend(out, (B, self.cfg.num_answers), "logits"))
        self._sync()                                               # TensorFlow's producers have finished
        self._check(self.lib.regat_engine_forward_dl(self.h, *[b.ptr for b in lent], None))
        self._sync()                                               # the result is complete before TensorFlow reads it
        return out

    def train_step(self, visual, bb, q_att, q_last, target, lr, step=None):
        """train.py:103-113 for the hot path's variables: loss = mean BCE * num_answers, gradients, per-tensor clip_by_norm(0.25),
        Adamax.  `step` is 1-based (Adamax bias correction); defaults to one more than the last call.  Returns (loss, batch score)."""
        B, N, lent = self._inputs(visual, bb, q_att, q_last)
        lent.append(self._lend(target, (B, self.cfg.num_answers), "target"))
        self.steps_taken = int(step) if step is not None else self.steps_taken + 1
        self._sync()
        self._check(self.lib.regat_engine_train_step_dl(self.h, *[b.ptr for b in lent], float(lr), self.steps_taken, self._loss.data, None))
        host = np.zeros(2, dtype=np.float32)
        self._check(self.lib.regat_memcpy(host.ctypes.data, self._loss.data, 8, 2, None))        # returns after the step has completed
        return float(host[0]), float(host[1])
