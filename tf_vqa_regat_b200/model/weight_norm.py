"""WeightNorm wrapper -- mirrors model/weight_norm.py:9-49 of the reference.

W = l2_normalize(v, axis=None) * g: WHOLE-TENSOR (Frobenius) norm with a SCALAR g (weight_norm.py:27-29,36,41), so
W = alpha * v with alpha = g / ||v||; the GEMM reads v and applies alpha (and the bias) in its epilogue.  Variables, in
Keras order: v (kernel shape), g (scalar), bias (weight_norm.py:21-33 un-tracks the wrapped kernel).  Built lazily on
the first call: v Glorot-uniform, g = ||v||, bias zeros (weight_norm.py:25,35-37)."""
import numpy as np
import torch

from .. import _lib
from . import _rt


class Layer:
    """Minimal stand-in for tf.keras.layers.Layer: callable, tracks sub-layers in attribute order."""

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    def sublayers(self):
        out = []
        for v in self.__dict__.values():
            if isinstance(v, Layer):
                out.append(v)
            elif isinstance(v, (list, tuple)):
                out.extend(x for x in v if isinstance(x, Layer))
        return out

    @property
    def weights(self):
        """[(name, tensor)] own variables first, then children in attribute-assignment order (Keras-2 rule)."""
        out = list(getattr(self, "_own_weights", lambda: [])())
        for s in self.sublayers():
            out.extend(s.weights)
        return out

    def get_weights(self):
        return [t.detach().cpu().numpy() for _, t in self.weights]

    def set_weights(self, arrays):
        ws = self.weights
        if len(ws) != len(arrays):
            raise ValueError(f"set_weights: expected {len(ws)} arrays, got {len(arrays)}")
        for (name, t), a in zip(ws, arrays):
            a = np.asarray(a, dtype=np.float32)
            if tuple(a.shape) != tuple(t.shape):
                raise ValueError(f"set_weights: {name} has shape {tuple(t.shape)}, got {a.shape}")
            t.copy_(torch.from_numpy(a).to(t.device))


class Dropout(Layer):
    """Identity: the reference never passes training=True (train.py:104), so Keras Dropout is inert.
    Kept so that layer indices (classifier.layers.0 / .3) and constructor arguments match."""

    def __init__(self, rate):
        self.rate = rate

    def call(self, x):
        return x


class Activation(Layer):
    def __init__(self, name):
        if name != "relu":
            raise NotImplementedError(f"Activation('{name}'): only 'relu' occurs on the hot path")
        self.name = name


class Dense:
    """Descriptor standing in for tf.keras.layers.Dense(units, use_bias, activation=None)."""

    def __init__(self, units, input_shape=None, use_bias=True, activation=None):
        if activation is not None:
            raise NotImplementedError("Dense(activation=...) is not used by the reference's hot path")
        self.units, self.use_bias = units, use_bias


class Conv2D:
    """Descriptor standing in for tf.keras.layers.Conv2D(filters, kernel_size=(1,1), groups=G) (graph_att_layer.py:32-36)."""

    def __init__(self, filters, input_shape=None, kernel_size=(1, 1), groups=1):
        if tuple(kernel_size) != (1, 1):
            raise NotImplementedError("only 1x1 convolutions occur on the hot path")
        self.filters, self.groups, self.use_bias = filters, groups, True


class WeightNorm(Layer):
    def __init__(self, layer, **kwargs):
        if type(layer).__name__ not in ("Dense", "Conv2D"):
            raise ValueError("WeightNorm is only implemented with Dense and Conv2D layers.")   # weight_norm.py:12-13
        self.layer = layer
        self.built = False
        self.fused_relu = False        # set by FullyConnected when the next layer is Activation('relu')

    # ---- variables
    def build(self, in_dim, device):
        L = self.layer
        if isinstance(L, Dense):
            kshape, fan_in, fan_out = (in_dim, L.units), in_dim, L.units
        else:
            cin_g = in_dim // L.groups
            kshape, fan_in, fan_out = (1, 1, cin_g, L.filters), cin_g, L.filters
        n = int(np.prod(kshape))
        cols = kshape[-1]
        r64 = lambda x: (x + 63) // 64 * 64
        self._flat = torch.zeros(r64(n) + 64 + r64(cols) + 128, dtype=torch.float32, device=device)
        lim = np.sqrt(6.0 / (fan_in + fan_out))                         # Keras glorot_uniform
        v0 = _rt.next_rng().uniform(-lim, lim, n).astype(np.float32)
        self.v = self._flat[:n].view(kshape)
        self.g = self._flat[r64(n):r64(n) + 1].view(())
        self.bias = self._flat[r64(n) + 64:r64(n) + 64 + cols] if L.use_bias else None
        self.v.copy_(torch.from_numpy(v0).view(kshape))
        self.g.copy_(torch.linalg.vector_norm(self.v))                  # _init_norm, weight_norm.py:35-37
        self._g_off = r64(n)
        self._stats = self._flat[r64(n) + 64 + r64(cols):]              # [0]=sumsq [32]=alpha [64]=inv_norm
        self.built = True

    def rebind(self, flat, v_off, g_off, b_off):
        """Moves the variables into `flat` (a 1-D fp32 device buffer with the engine's layout: tensors at 64-element aligned
        offsets), keeping their values: from now on this layer and the engine that owns `flat` share one set of weights."""
        v, g, bias = self.v, self.g, self.bias
        stats = torch.zeros(128, dtype=torch.float32, device=flat.device)
        self._flat = flat[v_off:]
        self.v = flat[v_off:v_off + v.numel()].view(v.shape)
        self.g = flat[g_off:g_off + 1].view(())
        self.v.copy_(v); self.g.copy_(g)
        if bias is not None:
            self.bias = flat[b_off:b_off + bias.numel()]
            self.bias.copy_(bias)
        self._g_off = g_off - v_off
        self._stats = stats

    def _own_weights(self):
        if not self.built:
            return []
        w = [("v", self.v), ("g", self.g)]
        if self.bias is not None:
            w.append(("bias", self.bias))
        return w

    def alpha_ptr(self):
        """Recomputes alpha = g/||v|| on the device (weight_norm.py:46 recomputes the kernel every call)."""
        import ctypes as C
        l = _lib.lib()
        s = self._stats
        s[0:1].zero_()
        off = np.array([0], dtype=np.int64); numel = np.array([self.v.numel()], dtype=np.int64)
        cols = np.array([self.v.shape[-1]], dtype=np.int32); goff = np.array([self._g_off], dtype=np.int64)
        _lib.check(l.regat_wn_prepare(self._flat.data_ptr(), off.ctypes.data, numel.ctypes.data, cols.ctypes.data, 1, s.data_ptr(),
                                      None, None, None, _rt.stream()))
        _lib.check(l.regat_wn_alpha(self._flat.data_ptr(), goff.ctypes.data, 1, s.data_ptr(), s[32:].data_ptr(), s[64:].data_ptr(),
                                    _rt.stream()))
        return s[32:].data_ptr()

    # ---- call
    def call(self, inputs, out=None, out_ld=None, relu=None):
        """Dense: [..., in] -> [..., units].  Conv2D: [R,1,1,G*cin] -> [R,1,1,filters] (NHWC, 1x1, grouped).
        out / out_ld let a caller place the result in a column slice of a wider buffer."""
        x = _rt.need_cuda(inputs, "WeightNorm input")
        in_dim = x.shape[-1]
        if not self.built:
            self.build(in_dim, x.device)
        rows = x.numel() // in_dim
        relu = self.fused_relu if relu is None else relu
        epi = _lib.Epilogue()
        epi.alpha = self.alpha_ptr()
        epi.bias = self.bias.data_ptr() if self.bias is not None else None
        epi.relu = int(bool(relu))
        L = self.layer
        if isinstance(L, Dense):
            if in_dim != self.v.shape[0]:
                raise ValueError(f"Dense kernel expects last dim {self.v.shape[0]}, got {in_dim}")
            y = out if out is not None else _rt.empty(*x.shape[:-1], L.units, device=x.device)
            ldc = out_ld if out_ld is not None else L.units
            _rt.gemm(0, 0, rows, L.units, in_dim, x.data_ptr(), in_dim, self.v.data_ptr(), L.units,
                     y.data_ptr() if isinstance(y, torch.Tensor) else y, ldc, epi)
            return y
        # grouped 1x1 conv: group g maps input channels [g*cin, (g+1)*cin) to output channels [g*cout, (g+1)*cout)
        G, cin, F = L.groups, in_dim // L.groups, L.filters
        cout = F // G
        y = _rt.empty(rows, 1, 1, F, device=x.device)
        for gi in range(G):
            e2 = _lib.Epilogue()
            e2.alpha, e2.relu = epi.alpha, epi.relu
            e2.bias = (self.bias.data_ptr() + 4 * gi * cout) if self.bias is not None else None
            _rt.gemm(0, 0, rows, cout, cin, x.data_ptr() + 4 * gi * cin, in_dim, self.v.data_ptr() + 4 * gi * cout, F,
                     y.data_ptr() + 4 * gi * cout, F, e2)
        return y
