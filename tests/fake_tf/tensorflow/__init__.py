"""TEST INFRASTRUCTURE: a stand-in for the handful of TensorFlow names integration/regat_b200_tf.py touches, backed by torch
(TensorFlow is not installable in this image).  It hands out real DLPack capsules of real device buffers, so the binding's
struct parsing, pointer plumbing and call sequence run exactly as they would against tf.experimental.dlpack."""
import contextlib

import torch

from . import experimental  # noqa: F401

float32 = torch.float32
int32 = torch.int32


class _Dev:
    name = None


_current = ["cpu"]


@contextlib.contextmanager
def device(name):
    prev = _current[0]
    _current[0] = "cuda:0" if "GPU" in name.upper() and torch.cuda.is_available() else "cpu"
    try:
        yield
    finally:
        _current[0] = prev


def zeros(shape, dtype=float32):
    return torch.zeros(*[int(s) for s in shape], dtype=dtype, device=_current[0])


def constant(value, dtype=float32):
    return torch.as_tensor(value, dtype=dtype).to(_current[0])
