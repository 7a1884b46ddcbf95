"""tf.keras.initializers stand-in (by name).  Values only matter until the golden generator
overwrites them with the seeded synthetic parameters; shapes and dtypes matter always."""
import math

import torch

from .._core import random as _random


def _fans(shape):
    if len(shape) < 1:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = 1
    for s in shape[:-2]:
        rf *= s
    return shape[-2] * rf, shape[-1] * rf


def get(name):
    if callable(name):
        return name
    if name == "zeros":
        return lambda shape, dt: torch.zeros(shape, dtype=dt)
    if name == "ones":
        return lambda shape, dt: torch.ones(shape, dtype=dt)
    if name == "random_normal":
        return lambda shape, dt: (0.05 * torch.randn(shape, generator=_random._gen, dtype=torch.float64)).to(dt)
    if name == "glorot_uniform":
        def init(shape, dt):
            fi, fo = _fans(shape)
            lim = math.sqrt(6.0 / (fi + fo))
            return ((torch.rand(shape, generator=_random._gen, dtype=torch.float64) * 2 - 1) * lim).to(dt)
        return init
    if name == "orthogonal":
        def init(shape, dt):
            a = torch.randn(shape, generator=_random._gen, dtype=torch.float64)
            q, r = torch.linalg.qr(a if shape[0] >= shape[1] else a.T)
            q = q * torch.sign(torch.diagonal(r))
            return (q if shape[0] >= shape[1] else q.T).to(dt)
        return init
    raise ValueError(f"initializer {name!r} is not part of the stand-in")
