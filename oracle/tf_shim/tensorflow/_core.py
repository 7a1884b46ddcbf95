"""Core of the TensorFlow stand-in (see oracle/tf_shim/README.md): tensors, variables, the
ops the reference's hot-path files call, GradientTape.  torch-CPU underneath.

TEST INFRASTRUCTURE.  Each primitive restates the published TensorFlow semantics named in
its docstring and nothing else; all model logic stays in the reference's own files, which
are executed unmodified on top of this.
"""
import numpy as np
import torch

newaxis = None


# ----------------------------------------------------------------------------- dtypes
class DType:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"tf.{self.name}"


float32 = DType("float32")
float64 = DType("float64")
int32 = DType("int32")
int64 = DType("int64")
bool_ = DType("bool")

_STATE = {"floatx": torch.float32}


def set_floatx(name):
    """tf.keras.backend.set_floatx (the reference carries a commented-out call with 'float64',
    main.py:101).  Under floatx=float64 the shim also maps tf.float32 to float64, so that the
    reference's explicit tf.float32 constants (graph_att_layer.py:66,85,92; relation_encoder.py:21)
    follow the run's working precision instead of raising dtype mismatches."""
    _STATE["floatx"] = {"float32": torch.float32, "float64": torch.float64}[name]


def floatx():
    return _STATE["floatx"]


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    name = dtype.name if isinstance(dtype, DType) else str(dtype)
    if name in ("float32", "float64"):
        return torch.float64 if (name == "float64" or _STATE["floatx"] == torch.float64) else torch.float32
    return {"int32": torch.int32, "int64": torch.int64, "bool": torch.bool}[name]


# ----------------------------------------------------------------------------- tensors
class Tensor(torch.Tensor):
    """An immutable-looking tensor: augmented assignment rebinds (TF tensors have no in-place
    ops; relation_encoder.py:89 `visual += imp_rel` relies on that)."""

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy()

    def __format__(self, spec):
        # EagerTensor formats through NumPy (train.py:133 prints a scalar tensor with :.4f)
        return format(self.numpy().item() if self.dim() == 0 else self.numpy(), spec)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        # results are plain shim Tensors whatever mix of Tensor / Variable went in
        with torch._C.DisableTorchFunctionSubclass():
            ret = func(*args, **(kwargs or {}))
        return _retype(ret)


class Variable(Tensor):
    """tf.Variable: a leaf that GradientTape watches; assign() overwrites the value."""

    @staticmethod
    def make(value, name=None, trainable=True):
        data = torch.as_tensor(value).detach().clone()
        v = torch.Tensor._make_subclass(Variable, data, bool(trainable) and data.is_floating_point())
        v.var_name = name
        v.trainable = trainable
        return v

    def assign(self, value):
        with torch.no_grad():
            self.copy_(torch.as_tensor(value, dtype=self.dtype))
        return self

    def assign_add(self, value):
        with torch.no_grad():
            self.add_(torch.as_tensor(value, dtype=self.dtype))
        return self

    def assign_sub(self, value):
        with torch.no_grad():
            self.sub_(torch.as_tensor(value, dtype=self.dtype))
        return self

    def read_value(self):
        return _w(self.detach().clone())


def _retype(r):
    if isinstance(r, torch.Tensor):
        return r if type(r) is Tensor else r.as_subclass(Tensor)
    if isinstance(r, (tuple, list)) and type(r) in (tuple, list):
        return type(r)(_retype(e) for e in r)
    return r


def _w(x):
    return x.as_subclass(Tensor) if isinstance(x, torch.Tensor) else x


def _t(x, dtype=None):
    """convert_to_tensor: NumPy float arrays and Python floats become floatx (Keras casts
    layer inputs to the layer's compute dtype); integers and bools keep their kind."""
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    a = np.asarray(x)
    if dtype is None:
        dtype = _STATE["floatx"] if a.dtype.kind == "f" else None
    return torch.as_tensor(a).to(dtype) if dtype is not None else torch.as_tensor(a)


def convert_to_tensor(value, dtype=None):
    return _w(_t(value, _dt(dtype)))


def constant(value, dtype=None):
    return _w(_t(value, _dt(dtype)))


def identity(x):
    return _w(_t(x).clone())


def cast(x, dtype):
    return _w(_t(x).to(_dt(dtype)))


def shape(x):
    return _w(torch.tensor(list(_t(x).shape), dtype=torch.int32))


def ones(shape, dtype=float32):
    return _w(torch.ones(tuple(int(s) for s in shape), dtype=_dt(dtype)))


def zeros(shape, dtype=float32):
    return _w(torch.zeros(tuple(int(s) for s in shape), dtype=_dt(dtype)))


def ones_like(x, dtype=None):
    return _w(torch.ones_like(_t(x), dtype=_dt(dtype)))


def reshape(x, shape):
    return _w(_t(x).reshape(tuple(int(s) for s in shape)))


def transpose(x, perm=None):
    x = _t(x)
    return _w(x.permute(*(perm if perm is not None else reversed(range(x.dim())))))


def expand_dims(x, axis):
    return _w(_t(x).unsqueeze(axis))


def squeeze(x, axis=None):
    return _w(_t(x).squeeze() if axis is None else _t(x).squeeze(axis))


def broadcast_to(x, shape):
    return _w(_t(x).expand(tuple(int(s) for s in shape)))


def tile(x, multiples):
    return _w(_t(x).repeat(*[int(m) for m in multiples]))


def concat(values, axis):
    return _w(torch.cat([_t(v) for v in values], dim=axis))


def stack(values, axis=0):
    return _w(torch.stack([_t(v) for v in values], dim=axis))


def matmul(a, b):
    """tf.matmul: batched over leading dims, contracts last of a with second-to-last of b."""
    return _w(torch.matmul(_t(a), _t(b)))


def sqrt(x):
    return _w(torch.sqrt(_t(x)))


def square(x):
    return _w(torch.square(_t(x)))


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return _w(x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims))


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return _w(x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims))


def not_equal(a, b):
    return _w(torch.ne(_t(a), _t(b)))


def where(condition, x=None, y=None):
    """tf.where: one argument -> int64 coordinates [n, rank] of the true elements; three -> select."""
    c = _t(condition)
    if x is None and y is None:
        return _w(torch.nonzero(c))
    return _w(torch.where(c, _t(x), _t(y)))


class IndexedSlices:
    """tf.IndexedSlices: what GradientTape returns for a variable that was read through tf.nn.embedding_lookup / tf.gather
    (array_grad.py, _GatherV2Grad / _ResourceGatherGrad): `values[k]` is the gradient of occurrence k of row `indices[k]`,
    duplicates NOT summed."""

    def __init__(self, values, indices, dense_shape):
        self.values, self.indices, self.dense_shape = values, indices, tuple(dense_shape)

    def to_dense(self):
        v = _t(self.values)
        out = torch.zeros(self.dense_shape, dtype=v.dtype)
        out.index_add_(0, _t(self.indices).long(), v)
        return _w(out)

    def numpy(self):
        return self.to_dense().numpy()


def clip_by_norm(t, clip_norm):
    """tf.clip_by_norm (clip_ops.py): t * clip_norm / max(||t||_2, clip_norm), the norm taken
    over the whole tensor, with the all-zero tensor mapped to itself.  For IndexedSlices the norm is that of `.values`
    -- the un-deduplicated occurrences -- and the result is IndexedSlices again (clip_ops.py: `values = t.values if
    isinstance(t, IndexedSlices) else t` ... `return IndexedSlices(values_clip, t.indices, t.dense_shape)`)."""
    if isinstance(t, IndexedSlices):
        return IndexedSlices(clip_by_norm(t.values, clip_norm), t.indices, t.dense_shape)
    t = _t(t)
    l2sum = (t * t).sum()
    pred = l2sum > 0
    l2sum_safe = torch.where(pred, l2sum, torch.ones_like(l2sum))
    l2norm = torch.where(pred, torch.sqrt(l2sum_safe), l2sum)
    return _w((t * clip_norm) / torch.maximum(l2norm, torch.as_tensor(clip_norm, dtype=t.dtype)))


def function(fn=None, **_):
    """tf.function: tracing changes no arithmetic; the Python body runs eagerly."""
    return fn if fn is not None else (lambda f: f)


class GradientTape:
    """tf.GradientTape over trainable variables: gradient() is reverse-mode autodiff of a scalar."""
    last = None                                    # (sources, gradients) of the most recent gradient() call

    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    _sparse = {}                                   # id(variable) -> [(indices, values)] recorded by embedding_lookup's backward

    def gradient(self, target, sources):
        sources = list(sources)
        GradientTape._sparse = {}
        g = torch.autograd.grad(_t(target), sources, allow_unused=True, retain_graph=self.persistent)
        out = []
        for src, x in zip(sources, g):
            rec = GradientTape._sparse.get(id(src))
            if x is None:
                out.append(None)
            elif rec:                              # the variable was only read through lookups: TensorFlow hands back IndexedSlices
                out.append(IndexedSlices(_w(torch.cat([v for _, v in rec], 0)), _w(torch.cat([i for i, _ in rec], 0)), src.shape))
            else:
                out.append(_w(x))
        GradientTape.last = (sources, out)
        return out


class _Lookup(torch.autograd.Function):
    """params[ids] whose backward also records the sparse form of the gradient (indices, per-occurrence values)."""

    @staticmethod
    def forward(ctx, params, ids, key):
        ctx.ids, ctx.key, ctx.pshape = ids, key, params.shape
        return params[ids]

    @staticmethod
    def backward(ctx, g):
        ids = ctx.ids.reshape(-1)
        vals = g.reshape(ids.numel(), -1)
        GradientTape._sparse.setdefault(ctx.key, []).append((ids, vals))
        dense = torch.zeros(ctx.pshape, dtype=g.dtype)
        dense.index_add_(0, ids, vals)
        return dense, None, None


# ----------------------------------------------------------------------------- tf.nn / tf.math
class nn:
    @staticmethod
    def relu(x):
        return _w(torch.relu(_t(x)))

    @staticmethod
    def softmax(x, axis=-1):
        return _w(torch.softmax(_t(x), dim=axis))

    @staticmethod
    def l2_normalize(x, axis=None, epsilon=1e-12):
        """x * rsqrt(max(sum(x**2, axis), epsilon)); axis=None reduces over every element."""
        x = _t(x)
        ss = (x * x).sum() if axis is None else (x * x).sum(dim=axis, keepdim=True)
        return _w(x * torch.rsqrt(torch.clamp(ss, min=epsilon)))

    @staticmethod
    def sigmoid_cross_entropy_with_logits(labels=None, logits=None):
        """max(x, 0) - x * z + log(1 + exp(-|x|)), element-wise."""
        x, z = _t(logits), _t(labels)
        return _w(torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x))))

    @staticmethod
    def embedding_lookup(params, ids):
        if isinstance(params, Variable):           # gradient w.r.t. the table comes back as IndexedSlices (see GradientTape.gradient)
            return _w(_Lookup.apply(params, _t(ids).long(), id(params)))
        return _w(_t(params)[_t(ids).long()])


class math:
    sqrt = staticmethod(sqrt)

    @staticmethod
    def maximum(a, b):
        return _w(torch.maximum(_t(a), _t(b)))

    @staticmethod
    def log(x):
        return _w(torch.log(_t(x)))


class random:
    _gen = torch.Generator().manual_seed(0)

    @staticmethod
    def set_seed(seed):
        random._gen.manual_seed(int(seed))


class _Logger:
    def setLevel(self, *_):
        pass


def get_logger():
    return _Logger()
