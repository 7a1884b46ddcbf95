"""ctypes binding of libregat.so (include/regat.h).  Fails loudly: if the shared library is missing
or there is no CUDA device there is NO fallback path -- every op raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REGAT_LIB") or os.path.join(_HERE, "libregat.so")     # REGAT_LIB: same-session A/B of two builds (tools/ab_lib.sh)

F32, BF16 = 0, 1


class RegatError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libregat status {status}: {msg}")
        self.status = status


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("v_dim", "q_dim", "rel_dim", "num_heads", "pos_emb_dim", "nongt_dim", "dir_num",
                                         "num_answers", "label_bias", "residual")] + \
               [(n, C.c_float) for n in ("grad_clip", "beta1", "beta2", "eps")]


class Epilogue(C.Structure):
    _fields_ = [("alpha", C.c_void_p), ("alpha_cols", C.c_int32), ("bias", C.c_void_p),
                ("addend", C.c_void_p), ("addend_ld", C.c_int32), ("addend_rows", C.c_int32), ("row_scale", C.c_void_p),
                ("relu", C.c_int32), ("accumulate", C.c_int32), ("gate", C.c_void_p), ("gate_ld", C.c_int32),
                ("c2", C.c_void_p), ("c2_ld", C.c_int32), ("c2_rows_in", C.c_int32), ("c2_rows_keep", C.c_int32),
                ("split_k", C.c_int32)]


GRAD_READY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64)
_lib = None
vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes; every function returns int.  Kept in the order of include/regat.h.
SIGNATURES = {
    "regat_abi_version": [],
    "regat_last_error": [C.c_char_p, C.c_size_t],
    "regat_default_config": [C.POINTER(Config)],
    "regat_device_count": [],
    "regat_memcpy": [vp, vp, i64, i32, vp],
    "regat_device_synchronize": [],
    "regat_position_embedding": [vp, i32, i32, i32, i32, vp, vp, vp],
    "regat_wn_prepare": [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "regat_wn_alpha": [vp, vp, i32, vp, vp, vp, vp],
    "regat_gemm_trace": [vp],
    "regat_gemm": [i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, i32, C.POINTER(Epilogue), vp],
    "regat_geoattn_fwd": [i32] * 8 + [vp, vp, vp, vp, vp, vp, i64, vp, vp, i64, vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "regat_attn_bwd": [i32] * 7 + [vp] * 9,
    "regat_geo_bwd": [i32] * 6 + [vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp],
    "regat_geo_bwd_ex": [i32] * 6 + [vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, i32, vp],
    "regat_geoattn_fast_supported": [i32, i32],
    "regat_geoattn_fwd_fast": [i32] * 7 + [vp, vp, vp, vp, vp, i64, vp, vp, i64, vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "regat_attn_bwd_fast": [i32] * 6 + [vp] * 11,
    "regat_geo_bwd_fast": [i32] * 6 + [vp, vp, vp, vp, i64, vp, i64, vp],
    "regat_explicit_pair_bias": [i32] * 5 + [vp, vp, vp, vp, vp],
    "regat_explicit_pair_bias_bwd": [i32] * 6 + [vp, vp, vp, vp, vp],
    "regat_graphattn_explicit_fwd": [i32] * 7 + [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp],
    "regat_graphattn_explicit_bwd": [i32] * 7 + [vp] * 10,
    "regat_butd_pool_fwd": [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "regat_butd_pool_bwd": [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp],
    "regat_pad_ragged": [i32, i32, i32, i64, vp, vp, vp, vp, vp],
    "regat_cast": [i32, i32, vp, vp, i64, vp],
    "regat_dp_reduce_bcast": [vp, C.c_uint64, vp, i32, i32, i64, i64, C.c_uint32, i32, vp],
    "regat_dp_allreduce_f32": [vp, C.c_uint64, vp, i32, i32, i64, i64, C.c_uint32, i32, vp],
    "regat_dp_allreduce_f32_dev": [vp, C.c_uint64, vp, i32, i32, i64, i64, vp, i32, vp],
    "regat_dp_wait_unpack": [vp, vp, vp, i32, i32, i64, i64, C.c_uint32, vp],
    "regat_concat_visual_question": [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    "regat_butd_prep": [i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp],
    "regat_mul": [i32, i32, i32, vp, i32, vp, i32, vp, i32, vp],
    "regat_bce_fwd_bwd": [i32, i32, vp, i32, vp, vp, vp, vp, i32, i32, vp],
    "regat_engine_create": [C.POINTER(Config), i32, i32, i32, C.POINTER(vp)],
    "regat_engine_destroy": [vp],
    "regat_engine_sizes": [vp, C.POINTER(i64), C.POINTER(i64)],
    "regat_engine_param": [vp, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "regat_engine_bind": [vp, vp, vp, vp, vp, vp, i64],
    "regat_engine_forward": [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp],
    "regat_engine_fwd_bwd": [vp, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp],
    "regat_engine_update": [vp, f32, i32, vp],
    "regat_engine_train_step": [vp, i32, i32, vp, vp, vp, vp, vp, f32, i32, vp, vp],
    "regat_engine_set_lr": [vp, f32, vp],
    "regat_engine_set_step": [vp, i32, vp],
    "regat_engine_get_step": [vp, C.POINTER(C.c_int), C.POINTER(C.c_float), vp],
    "regat_engine_train_step_dev": [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp],
    "regat_engine_set_dp": [vp, vp, C.c_uint64, vp, i32, i32, i32],
    "regat_engine_refresh_weights": [vp, vp],
    "regat_engine_profile": [vp, i32],
    "regat_engine_profile_read": [vp, i32, vp, vp, C.POINTER(C.c_int)],
    "regat_engine_set_grad_callback": [vp, vp, vp],
    "regat_engine_last_launches": [vp],
    "regat_engine_params_changed": [vp],
    "regat_engine_config": [vp, C.POINTER(Config)],
    "regat_engine_set_wave_div": [vp, vp],
    "regat_engine_finalize_grads": [vp, vp],
    "regat_engine_buffer": [vp, C.c_char_p, C.POINTER(vp)],
    "regat_engine_forward_dl": [vp, vp, vp, vp, vp, vp, vp],
    "regat_engine_train_step_dl": [vp, vp, vp, vp, vp, vp, f32, i32, vp, vp],
    "regat_q_embed_fwd": [vp, i64, i32, i32, vp, vp, vp, vp],
    "regat_q_embed_bwd": [vp, i64, i32, i32, i32, vp, vp, vp, vp],
    "regat_q_gru_gates_fwd": [i32, i32, vp, i64, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp],
    "regat_q_gru_gates_bwd": [i32, i32, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp],
    "regat_q_tanh_fwd": [vp, i64, vp],
    "regat_q_tanh_bwd": [vp, vp, i64, vp],
    "regat_q_batch_softmax_fwd": [vp, i32, i32, vp, vp],
    "regat_q_batch_softmax_bwd": [vp, vp, i32, i32, vp, vp],
    "regat_q_pool_fwd": [vp, vp, i32, i32, i32, vp, vp],
    "regat_q_pool_bwd": [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp],
    "regat_q_dot": [vp, vp, i64, vp, vp],
    "regat_q_wn_alpha": [vp, vp, vp, vp],
    "regat_q_wn_bwd": [vp, vp, vp, vp, vp, i64, vp, vp, vp],
    "regat_q_clip_adamax": [vp, vp, vp, vp, i64, vp, f32, f32, i32, f32, f32, f32, vp],
    "regat_q_embed_sumsq": [vp, i64, i32, i32, i32, i32, vp, vp, vp],
    "regat_q_embed_clip_adamax": [vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, f32, f32, i32, f32, f32, f32, vp],
}


def lib():
    """The loaded library.  Raises if it has not been built (python -m tf_vqa_regat_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RegatError(-6, f"{LIB_PATH} is missing: build it with `python -m tf_vqa_regat_b200.build` "
                                 "(there is no CPU / PyTorch fallback for this path)")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        if l.regat_abi_version() != 1:
            raise RegatError(-1, "libregat.so ABI version mismatch")
        _lib = l
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    lib().regat_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(status):
    if status != 0:
        raise RegatError(status, last_error())


def wave_divisors(feat_dim=64, wave_length=1000):
    """fp32 divisors 1000^(8k/feat_dim), evaluated the way position_emb.py:98-100 does (np.power in float32)."""
    k = np.arange(0, feat_dim / 8, dtype=np.float32)
    return np.ascontiguousarray(np.power(np.full((1,), wave_length, dtype=np.float32), (8.0 / feat_dim) * k), dtype=np.float32)


def ptr(t):
    """Device (or host, for numpy) pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()
