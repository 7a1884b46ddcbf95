"""Runtime helpers shared by the layer mirror: torch supplies device memory and the stream, libregat.so does the math."""
import ctypes as C

import numpy as np
import torch

from .. import _lib

DT = _lib.F32                      # the layer mirror runs the exact-fp32 kernels (parity surface); bf16 lives in engine.py
_seed = [20261018]


def set_seed(seed: int):
    """Seed of the Glorot initialisers (the reference seeds TF in main.py:106-108)."""
    _seed[0] = int(seed)


def next_rng():
    _seed[0] += 1
    return np.random.default_rng(_seed[0])


def stream():
    return torch.cuda.current_stream().cuda_stream


def need_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.RegatError(-4, f"{what}: expected a CUDA torch tensor (there is no CPU path); got {type(t).__name__}"
                                  f"{'' if not isinstance(t, torch.Tensor) else ' on ' + str(t.device)}")
    if t.dtype != torch.float32:
        raise _lib.RegatError(-3, f"{what}: expected float32, got {t.dtype}")
    return t.contiguous()


def empty(*shape, device):
    return torch.empty(*shape, dtype=torch.float32, device=device)


def gemm(tA, tB, M, N, K, A, lda, B, ldb, Cp, ldc, epi=None):
    """A, B, Cp are integer device pointers (so that column / row slices can be addressed)."""
    _lib.check(_lib.lib().regat_gemm(DT, int(tA), int(tB), M, N, K, A, lda, B, ldb, Cp, ldc, DT,
                                     C.byref(epi) if epi is not None else None, stream()))
