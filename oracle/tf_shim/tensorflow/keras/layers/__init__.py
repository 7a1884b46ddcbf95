"""tf.keras.layers stand-in: the Layer protocol (build on first call, weight tracking in
Keras-2 order) and the stock layers the reference instantiates (five on the hot path, GRU in the
question front-end).  TEST INFRASTRUCTURE.

Keras-2 semantics restated here (keras/engine/base_layer.py):
  * __call__ converts inputs, builds the layer on first use (build(input_shape), then
    built=True whether or not the subclass called super().build), then runs call().
  * `training` is not forwarded unless the caller passes it; Dropout without it is the
    identity outside fit() (the learning phase is 0 in eager mode).
  * weights = the layer's own variables in creation order, then the variables of tracked
    sub-layers in attribute-assignment order; lists of layers are tracked element by element,
    including elements appended later.
  * assigning to an attribute that currently holds a tracked variable un-tracks that variable
    (Layer.__setattr__ deletes the old attribute first) -- weight_norm.py:31 relies on it.
"""
import numpy as np
import torch

from .. import initializers as _init
from ..._core import Variable, _dt, _t, _w, floatx


class Layer:
    def __init__(self, trainable=True, name=None, dtype=None, **kwargs):
        object.__setattr__(self, "_own_weights", [])
        self.trainable = trainable
        self.name = name or type(self).__name__.lower()
        self.built = False
        self.input_shape_arg = kwargs.pop("input_shape", None)

    # -- tracking ---------------------------------------------------------------------------
    def __setattr__(self, name, value):
        old = self.__dict__.get(name)
        if isinstance(old, Variable) and old is not value:
            self._own_weights[:] = [w for w in self._own_weights if w is not old]
            del self.__dict__[name]                    # re-insertion moves the attribute to the end, as in Keras
        object.__setattr__(self, name, value)

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, **_):
        dt = _dt(dtype) or floatx()
        init = _init.get(initializer if initializer is not None else "glorot_uniform")
        v = Variable.make(init(tuple(int(s) for s in shape), dt), name=name, trainable=trainable)
        self._own_weights.append(v)
        return v

    def _children(self):
        for k, v in self.__dict__.items():
            if k.startswith("_own"):
                continue
            if isinstance(v, Layer):
                yield v
            elif isinstance(v, (list, tuple)):
                for e in v:
                    if isinstance(e, Layer):
                        yield e

    @property
    def weights(self):
        out, seen = [], set()
        for w in list(self._own_weights) + [w for c in self._children() for w in c.weights]:
            if id(w) not in seen:
                seen.add(id(w))
                out.append(w)
        return out

    @property
    def trainable_weights(self):
        """Keras: a layer with trainable=False contributes none of its variables, nor do its sub-layers'."""
        if not self.trainable:
            return []
        out, seen = [], set()
        own = [w for w in self._own_weights if getattr(w, "trainable", True)]
        for w in own + [w for c in self._children() for w in c.trainable_weights]:
            if id(w) not in seen:
                seen.add(id(w))
                out.append(w)
        return out

    @property
    def trainable_variables(self):
        return self.trainable_weights

    @property
    def variables(self):
        return self.weights

    def named_weights(self, prefix=""):
        """(path, variable) in `weights` order; paths are attribute names joined by '.', list elements by index."""
        out = [(f"{prefix}/{w.var_name}", w) for w in self._own_weights]
        for k, v in self.__dict__.items():
            if k.startswith("_own"):
                continue
            if isinstance(v, Layer):
                out += v.named_weights(f"{prefix}.{k}" if prefix else k)
            elif isinstance(v, (list, tuple)):
                for i, e in enumerate(v):
                    if isinstance(e, Layer):
                        out += e.named_weights(f"{prefix}.{k}.{i}" if prefix else f"{k}.{i}")
        return out

    def set_weights(self, values):
        ws = self.weights
        assert len(ws) == len(values)
        for w, v in zip(ws, values):
            w.assign(v)

    def get_weights(self):
        return [w.numpy().copy() for w in self.weights]

    # -- call protocol ----------------------------------------------------------------------
    def build(self, input_shape):
        self.built = True

    def __call__(self, *args, **kwargs):
        args = tuple(_w(_t(a)) if isinstance(a, (np.ndarray, torch.Tensor)) else a for a in args)
        if not self.built:
            first = args[0] if args else None
            self.build(tuple(first.shape) if isinstance(first, torch.Tensor) else None)
            self.built = True
        return self.call(*args, **kwargs)


class Wrapper(Layer):
    """keras.layers.Wrapper: holds `layer`."""

    def __init__(self, layer, **kwargs):
        super().__init__(**kwargs)
        self.layer = layer


class Dense(Layer):
    """keras.layers.Dense: kernel [in, units] (Glorot uniform), bias [units] (zeros);
    outputs = inputs . kernel over the last axis (+ bias) (activation is None at every call site)."""

    def __init__(self, units, activation=None, use_bias=True, **kwargs):
        super().__init__(**kwargs)
        assert activation is None
        self.units, self.use_bias = int(units), use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", shape=[int(input_shape[-1]), self.units])
        self.bias = self.add_weight("bias", shape=[self.units], initializer="zeros") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        y = torch.matmul(_t(inputs), _t(self.kernel))
        return _w(y + self.bias if self.use_bias else y)


class Conv2D(Layer):
    """keras.layers.Conv2D, NHWC, stride 1, 'valid', 1x1 kernels only (the one call site,
    graph_att_layer.py:32-36).  kernel [kh, kw, Cin/groups, filters].  Grouped convolution:
    input channels and filters are both split into `groups` contiguous blocks; block g of the
    output sees block g of the input through kernel[..., g*filters/groups:(g+1)*filters/groups]."""

    def __init__(self, filters, kernel_size, groups=1, use_bias=True, **kwargs):
        super().__init__(**kwargs)
        self.filters, self.kernel_size, self.groups, self.use_bias = int(filters), tuple(kernel_size), int(groups), use_bias
        assert self.kernel_size == (1, 1)

    def build(self, input_shape):
        cin = int(input_shape[-1])
        assert cin % self.groups == 0 and self.filters % self.groups == 0
        self.kernel = self.add_weight("kernel", shape=list(self.kernel_size) + [cin // self.groups, self.filters])
        self.bias = self.add_weight("bias", shape=[self.filters], initializer="zeros") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        x = _t(inputs)
        b, h, w, cin = x.shape
        G = self.groups
        xg = x.reshape(b, h, w, G, cin // G)
        kg = _t(self.kernel)[0, 0].reshape(cin // G, G, self.filters // G)
        y = torch.einsum("bhwgc,cgo->bhwgo", xg, kg).reshape(b, h, w, self.filters)
        return _w(y + self.bias if self.use_bias else y)


class Dropout(Layer):
    calls_with_training = 0                            # how often any Dropout was asked to drop (stays 0: SURVEY A.2-Q1)

    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate

    def call(self, inputs, training=None):
        if training:
            Dropout.calls_with_training += 1
            raise NotImplementedError("the hot path never calls Dropout with training=True (train.py:104)")
        return inputs


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.fn = {"relu": torch.relu, "tanh": torch.tanh}[activation]

    def call(self, inputs):
        return _w(self.fn(_t(inputs)))


class GRU(Layer):
    """keras.layers.GRU (TF2 defaults: activation tanh, recurrent_activation sigmoid, use_bias, reset_after=True, zero
    initial state, no masking unless a mask is passed -- the reference passes none).  Variables in Keras order:
    kernel [in, 3u], recurrent_kernel [u, 3u], bias [2, 3u] (row 0 input bias, row 1 recurrent bias); gate blocks z | r | h.

        x_t W + b_0 -> (xz, xr, xh);  h_{t-1} U + b_1 -> (hz, hr, hh)
        z = sigmoid(xz + hz);  r = sigmoid(xr + hr);  c = tanh(xh + r * hh);  h_t = z * h_{t-1} + (1 - z) * c

    The question front-end (language_model.py:96-131) is the only user; dropout is forced to 0 there (:105)."""

    def __init__(self, units, dropout=0.0, return_sequences=False, return_state=False, **kwargs):
        super().__init__()
        assert not dropout, "the reference forces GRU dropout to 0 (language_model.py:105)"
        self.units, self.return_sequences, self.return_state = int(units), return_sequences, return_state

    def build(self, input_shape):
        u = self.units
        self.kernel = self.add_weight("kernel", shape=[int(input_shape[-1]), 3 * u])
        self.recurrent_kernel = self.add_weight("recurrent_kernel", shape=[u, 3 * u], initializer="orthogonal")
        self.bias = self.add_weight("bias", shape=[2, 3 * u], initializer="zeros")
        self.built = True

    def call(self, inputs):
        x = _t(inputs)
        B, T, _ = x.shape
        u = self.units
        W, U, b = _t(self.kernel), _t(self.recurrent_kernel), _t(self.bias)
        h = torch.zeros(B, u, dtype=x.dtype)
        seq = []
        for t in range(T):
            xi = torch.matmul(x[:, t], W) + b[0]
            hi = torch.matmul(h, U) + b[1]
            z = torch.sigmoid(xi[:, :u] + hi[:, :u])
            r = torch.sigmoid(xi[:, u:2 * u] + hi[:, u:2 * u])
            c = torch.tanh(xi[:, 2 * u:] + r * hi[:, 2 * u:])
            h = z * h + (1 - z) * c
            seq.append(h)
        out = _w(torch.stack(seq, dim=1)) if self.return_sequences else _w(h)
        return (out, _w(h)) if self.return_state else out
