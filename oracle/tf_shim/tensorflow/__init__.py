"""TEST INFRASTRUCTURE ONLY -- a torch-CPU stand-in for the parts of TensorFlow 2.12 / Keras 2.12 that the
reference's forward path calls, so that the UNMODIFIED reference sources under ``/root/reference/tfimm`` can be
executed on CPU and used to pin ``oracle/*.py`` (TensorFlow itself is not installable here).

What is emulated is third-party behaviour (TensorFlow/Keras, pinned at ``poetry.lock:1405-1406, 608-609`` of the
reference), restated from its published semantics:

* ``tf.keras.layers.Layer``: lazy ``build(input_shape)`` on first call, ``add_weight``, attribute tracking of
  sub-layers / lists / dicts / ``tf.Variable``s, and TF name scopes for variable names
  (``<model>/<layer>/.../kernel:0``; a ``tf.keras.Sequential`` builds its layers in a fresh graph, so their
  variables carry no outer scope -- the reason the reference writes full paths into those layer names,
  ``tfimm/architectures/resnet.py:299-330, 478-540``).
* the layer classes listed in SURVEY.md 2.2 (Dense, Conv2D incl. groups / "same" padding, DepthwiseConv2D, Conv1D,
  ZeroPadding1D/2D, LayerNormalization, BatchNormalization (inference), Activation, ReLU, Dropout (inference),
  pooling layers, Flatten) and the ``tf.*`` tensor functions the five in-scope families use.
* tensors are ``torch.Tensor`` subclasses so that the reference's ``x.shape.ndims`` / ``mask.get_shape()`` work.

Nothing under ``tensorflow-image-models_b200/`` may import this package; it is put on ``sys.path`` only by
``oracle/ref_runner.py`` (used by ``tests/`` and ``tools/make_golden.py``).
"""
import inspect
import math as _pymath
import re
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

__version__ = "2.12.0-shim"

# ----------------------------------------------------------------------------------------------- dtypes
float16, bfloat16, float32, float64 = torch.float16, torch.bfloat16, torch.float32, torch.float64
int32, int64, uint8 = torch.int32, torch.int64, torch.uint8
bool = torch.bool  # noqa: A001  (mirrors tf.bool)

_FLOATX = ["float32"]
_NAMES = {"float16": float16, "bfloat16": bfloat16, "float32": float32, "float64": float64, "int32": int32,
          "int64": int64, "uint8": uint8, "bool": torch.bool}


def _dtype(d):
    if d is None:
        return _NAMES[_FLOATX[0]]
    if isinstance(d, torch.dtype):
        return d
    if isinstance(d, str):
        return _NAMES[d]
    if isinstance(d, np.dtype) or (isinstance(d, type) and issubclass(d, np.generic)):
        return _NAMES[np.dtype(d).name]
    raise TypeError(f"unsupported dtype {d!r}")


class TensorShape(tuple):
    @property
    def ndims(self):
        return len(self)

    rank = ndims

    def as_list(self):
        return list(self)


class Tensor(torch.Tensor):
    """Eager tensor: a torch tensor with the few tf.Tensor accessors the reference uses."""

    @property
    def shape(self):
        return TensorShape(torch.Tensor.shape.__get__(self))

    def get_shape(self):
        return self.shape

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).cpu().numpy().copy()


def _t(x, dtype=None):
    """Anything -> shim Tensor (numpy float64 stays float64, python floats become floatx)."""
    if isinstance(x, torch.Tensor):
        out = x if isinstance(x, Tensor) else x.as_subclass(Tensor)
    elif isinstance(x, np.ndarray):
        out = torch.from_numpy(np.ascontiguousarray(x)).as_subclass(Tensor)
    elif isinstance(x, (float, int)) or (isinstance(x, (list, tuple)) and not any(
            isinstance(e, torch.Tensor) for e in x)):
        arr = np.asarray(x)
        if arr.dtype == np.float64:
            out = torch.tensor(arr, dtype=_dtype(None)).as_subclass(Tensor)
        else:
            out = torch.from_numpy(arr).as_subclass(Tensor)
    elif isinstance(x, (list, tuple)):
        out = torch.stack([_t(e) for e in x]).as_subclass(Tensor)
    else:
        raise TypeError(f"cannot convert {type(x)} to a tensor")
    if dtype is not None:
        out = out.to(_dtype(dtype))
    return out


def convert_to_tensor(value, dtype=None, **_):
    return _t(value, dtype)


constant = convert_to_tensor

# ----------------------------------------------------------------------------------------------- name scopes
_SCOPE = []
_UIDS = {}


class name_scope:
    def __init__(self, name):
        self.parts = [p for p in str(name).split("/") if p]

    def __enter__(self):
        _SCOPE.extend(self.parts)
        return self

    def __exit__(self, *exc):
        del _SCOPE[len(_SCOPE) - len(self.parts):]
        return False


class _fresh_graph:
    """Variables created inside carry no outer name scope (Keras functional / Sequential graph build)."""

    def __enter__(self):
        self.saved = list(_SCOPE)
        del _SCOPE[:]

    def __exit__(self, *exc):
        _SCOPE[:] = self.saved
        return False


def _scoped(name):
    return "/".join(_SCOPE + [name]) + ":0"


class Variable(Tensor):
    def __new__(cls, initial_value=None, trainable=True, name=None, dtype=None, **_):
        data = _t(initial_value, dtype).detach().clone()
        obj = torch.Tensor._make_subclass(cls, data, False)
        obj._vname = _scoped(name or "Variable")
        obj._trainable = trainable
        return obj

    def __init__(self, *a, **k):
        pass

    @property
    def name(self):
        return self._vname

    @property
    def trainable(self):
        return self._trainable

    def assign(self, value):
        v = _t(value).to(self.dtype)
        if tuple(v.shape) != tuple(self.shape):
            raise ValueError(f"Cannot assign value of shape {tuple(v.shape)} to {self.name} {tuple(self.shape)}")
        with torch.no_grad():
            self.as_subclass(torch.Tensor).copy_(v.as_subclass(torch.Tensor))
        return self

    def value(self):
        return self.as_subclass(Tensor)

    def __deepcopy__(self, memo):
        raise TypeError("tf.Variable is not deep-copied by the shim")


# ----------------------------------------------------------------------------------------------- tensor functions
def shape(x):  # noqa: F811  (tf.shape)
    return TensorShape(_t(x).shape) if not isinstance(x, (tuple, list)) else TensorShape((len(x),))


def rank(x):
    return _t(x).dim()


def reshape(tensor, shape, name=None):  # noqa: A002
    return _t(tensor).reshape(tuple(int(s) for s in shape))


def transpose(a, perm=None, **_):
    a = _t(a)
    return a.permute(*perm) if perm is not None else a.permute(*reversed(range(a.dim())))


def unstack(value, num=None, axis=0):
    if isinstance(value, (tuple, list)) and not isinstance(value, torch.Tensor):
        return tuple(value)
    return tuple(torch.unbind(_t(value), dim=axis))


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], dim=axis)


def expand_dims(input, axis):  # noqa: A002
    return _t(input).unsqueeze(axis)


def concat(values, axis):
    if all(isinstance(v, (tuple, list)) and not isinstance(v, torch.Tensor) for v in values):
        out = []
        for v in values:
            out.extend(v)
        return TensorShape(out)
    return torch.cat([_t(v) for v in values], dim=axis)


def zeros(shape, dtype=None):  # noqa: A002
    return torch.zeros(tuple(shape), dtype=_dtype(dtype)).as_subclass(Tensor)


def ones(shape, dtype=None):  # noqa: A002
    return torch.ones(tuple(shape), dtype=_dtype(dtype)).as_subclass(Tensor)


def cast(x, dtype):
    return _t(x).to(_dtype(dtype))


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (tuple, list)) else (axis,)


def reduce_mean(input_tensor, axis=None, keepdims=False):
    x = _t(input_tensor)
    return x.mean() if axis is None else x.mean(dim=_axes(axis), keepdim=keepdims)


def reduce_sum(input_tensor, axis=None, keepdims=False):
    x = _t(input_tensor)
    return x.sum() if axis is None else x.sum(dim=_axes(axis), keepdim=keepdims)


def reduce_variance(input_tensor, axis=None, keepdims=False):
    x = _t(input_tensor)
    return x.var(dim=_axes(axis), unbiased=False, keepdim=keepdims)


def where(condition, x=None, y=None):
    c = _t(condition)
    xs, ys = x, y
    ref = next((v for v in (x, y) if isinstance(v, torch.Tensor)), None)
    if not isinstance(xs, torch.Tensor):
        xs = torch.tensor(xs, dtype=ref.dtype if ref is not None else _dtype(None))
    if not isinstance(ys, torch.Tensor):
        ys = torch.tensor(ys, dtype=ref.dtype if ref is not None else _dtype(None))
    return torch.where(c, xs, ys).as_subclass(Tensor)


def tile(input, multiples):  # noqa: A002
    return _t(input).repeat(*[int(m) for m in multiples])


def split(value, num_or_size_splits, axis=0):
    v = _t(value)
    if isinstance(num_or_size_splits, int):
        return list(torch.chunk(v, num_or_size_splits, dim=axis))
    return list(torch.split(v, list(num_or_size_splits), dim=axis))


def roll(input, shift, axis):  # noqa: A002
    sh = tuple(shift) if isinstance(shift, (tuple, list)) else (shift,)
    ax = tuple(axis) if isinstance(axis, (tuple, list)) else (axis,)
    return torch.roll(_t(input), shifts=tuple(int(s) for s in sh), dims=ax)


def repeat(input, repeats, axis=None):  # noqa: A002
    return torch.repeat_interleave(_t(input), int(repeats), dim=axis)


def gather(params, indices, axis=0, **_):
    return torch.index_select(_t(params), axis, _t(indices).long().reshape(-1)).reshape(
        *tuple(_t(params).shape[:axis]), *tuple(_t(indices).shape), *tuple(_t(params).shape[axis + 1:]))


def floor(x):
    return torch.floor(_t(x))


def pad(tensor, paddings, mode="CONSTANT", constant_values=0):
    x = _t(tensor)
    pads = [tuple(int(v) for v in p) for p in paddings]
    if mode.upper() == "CONSTANT":
        flat = []
        for lo, hi in reversed(pads):
            flat += [lo, hi]
        return F.pad(x, flat, value=constant_values)
    if mode.upper() == "REFLECT":
        # mirror without repeating the edge sample (tf.pad REFLECT), dimension by dimension
        for d, (lo, hi) in enumerate(pads):
            if lo == 0 and hi == 0:
                continue
            n = x.shape[d]
            idx = list(range(lo, 0, -1)) + list(range(n)) + list(range(n - 2, n - 2 - hi, -1))
            x = torch.index_select(x, d, torch.tensor(idx))
        return x
    raise NotImplementedError(mode)


def _same_pads(size, k_eff, s):
    out = -(-size // s)
    total = max((out - 1) * s + k_eff - size, 0)
    return total // 2, total - total // 2


def _conv2d_nhwc(x, kernel_hwio, strides, padding, dilation=(1, 1), groups=1):
    """tf.nn.conv2d semantics on NHWC input / HWIO filter ("SAME" pads more at the bottom/right)."""
    x = _t(x)
    kh, kw = kernel_hwio.shape[0], kernel_hwio.shape[1]
    sh, sw = strides
    dh, dw = dilation
    if padding.upper() == "SAME":
        pt, pb = _same_pads(x.shape[1], dh * (kh - 1) + 1, sh)
        pl, pr = _same_pads(x.shape[2], dw * (kw - 1) + 1, sw)
        x = F.pad(x, (0, 0, pl, pr, pt, pb))
    elif padding.upper() != "VALID":
        raise ValueError(padding)
    w = _t(kernel_hwio).permute(3, 2, 0, 1)
    y = F.conv2d(x.permute(0, 3, 1, 2), w.to(x.dtype), None, stride=(sh, sw), dilation=(dh, dw), groups=groups)
    return y.permute(0, 2, 3, 1)


def _two(v):
    return (int(v), int(v)) if isinstance(v, (int, np.integer)) else tuple(int(a) for a in v)


# ----------------------------------------------------------------------------------------------- tf.nn / math / ...
nn = types.ModuleType("tensorflow.nn")
linalg = types.ModuleType("tensorflow.linalg")
math_ = types.ModuleType("tensorflow.math")
image = types.ModuleType("tensorflow.image")
random = types.ModuleType("tensorflow.random")
initializers = types.ModuleType("tensorflow.initializers")


def _softmax(logits, axis=-1):
    return torch.softmax(_t(logits), dim=axis)


def _moments(x, axes, keepdims=False):
    x = _t(x)
    return x.mean(dim=tuple(axes), keepdim=keepdims), x.var(dim=tuple(axes), unbiased=False, keepdim=keepdims)


def _batch_normalization(x, mean, variance, offset, scale, variance_epsilon):
    inv = torch.rsqrt(_t(variance) + variance_epsilon)
    if scale is not None:
        inv = inv * scale
    return _t(x) * inv + ((offset if offset is not None else 0.0) - _t(mean) * inv)


def _depthwise_conv2d(input, filter, strides, padding, **_):  # noqa: A002
    c = filter.shape[2]
    w = _t(filter).reshape(filter.shape[0], filter.shape[1], 1, c * filter.shape[3])
    return _conv2d_nhwc(input, w, (strides[1], strides[2]), padding, groups=c)


def _matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def _cubic_weights(out_size, in_size):
    """One axis of ResizeBicubic(half_pixel_centers=True) -- the op tf.image.resize(method="bicubic",
    antialias=False) dispatches to (tensorflow/core/kernels/image/resize_bicubic_op.cc, restated): Keys kernel
    a = -0.5 tabulated at 1024 steps (the fractional offset is rounded to the table grid), float32 source
    coordinate (o + 0.5) * scale - 0.5, taps outside the image get weight 0 and the rest are renormalised."""
    table = 1024
    a = -0.5

    def near(x):
        return ((a + 2) * x - (a + 3)) * x * x + 1

    def far(x):
        return ((a * x - 5 * a) * x + 8 * a) * x - 4 * a

    lut0 = np.array([near(i / table) for i in range(table + 1)], dtype=np.float32)
    lut1 = np.array([far(i / table + 1.0) for i in range(table + 1)], dtype=np.float32)
    scale = np.float32(in_size) / np.float32(out_size)
    W = np.zeros((out_size, in_size), dtype=np.float64)
    for o in range(out_size):
        src = np.float32(np.float32(o + 0.5) * scale) - np.float32(0.5)
        base = int(np.floor(src))
        off = int(np.rint(np.float32(src - np.float32(base)) * np.float32(table)))
        taps = ((base - 1, lut1[off]), (base, lut0[off]), (base + 1, lut0[table - off]), (base + 2, lut1[table - off]))
        wsum = np.float32(0)
        kept = []
        for idx, wgt in taps:
            if 0 <= idx < in_size:
                kept.append((idx, wgt))
                wsum = np.float32(wsum + wgt)
        for idx, wgt in kept:
            W[o, idx] += float(np.float32(wgt / wsum))
    return W


def _resize(images, size, method="bilinear", **_):
    if method != "bicubic":
        raise NotImplementedError(method)
    x = _t(images)
    Wh = torch.from_numpy(_cubic_weights(int(size[0]), x.shape[1])).to(torch.float64)
    Ww = torch.from_numpy(_cubic_weights(int(size[1]), x.shape[2])).to(torch.float64)
    y = torch.einsum("oh,bhwc->bowc", Wh, x.to(torch.float64).as_subclass(torch.Tensor))
    y = torch.einsum("pw,bowc->bopc", Ww, y)
    return y.to(torch.float32).as_subclass(Tensor)  # tf.image.resize returns float32


_GEN = torch.Generator().manual_seed(0)


def _uniform(shape, minval=0.0, maxval=1.0, dtype=None, seed=None):  # noqa: A002
    return (torch.rand(tuple(shape), generator=_GEN, dtype=_dtype(dtype)) * (maxval - minval) + minval
            ).as_subclass(Tensor)


def _normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):  # noqa: A002
    return (torch.randn(tuple(shape), generator=_GEN, dtype=_dtype(dtype)) * stddev + mean).as_subclass(Tensor)


nn.softmax, nn.moments, nn.batch_normalization, nn.depthwise_conv2d = (
    _softmax, _moments, _batch_normalization, _depthwise_conv2d)
linalg.matmul = _matmul
matmul = _matmul
math_.sqrt = lambda x: torch.sqrt(_t(x))
math_.reduce_variance, math_.reduce_mean = reduce_variance, reduce_mean
math_.divide = lambda x, y: _t(x) / y
image.resize = _resize
random.uniform, random.normal = _uniform, _normal
math = math_  # exported as tf.math (the stdlib module is _pymath here)


# ----------------------------------------------------------------------------------------------- initializers
class Initializer:
    def __call__(self, shape, dtype=None, **kwargs):  # noqa: A002
        raise NotImplementedError

    def get_config(self):
        return {}


class Zeros(Initializer):
    def __call__(self, shape, dtype=None, **kwargs):  # noqa: A002
        return zeros(shape, dtype)


class Ones(Initializer):
    def __call__(self, shape, dtype=None, **kwargs):  # noqa: A002
        return ones(shape, dtype)


class Constant(Initializer):
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape, dtype=None, **kwargs):  # noqa: A002
        v = _t(np.asarray(self.value, dtype=np.float64)).to(_dtype(dtype))
        return zeros(shape, dtype) + (v.reshape(tuple(shape)) if v.numel() > 1 else v.reshape(()))


class RandomNormal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape, dtype=None, **kwargs):  # noqa: A002
        return _normal(shape, self.mean, self.stddev, dtype)


class TruncatedNormal(RandomNormal):
    def __call__(self, shape, dtype=None, **kwargs):  # noqa: A002
        x = torch.empty(tuple(shape), dtype=_dtype(dtype))
        torch.nn.init.trunc_normal_(x, self.mean, self.stddev, self.mean - 2 * self.stddev,
                                    self.mean + 2 * self.stddev, generator=_GEN)
        return x.as_subclass(Tensor)


class GlorotUniform(Initializer):
    def __call__(self, shape, dtype=None, **kwargs):  # noqa: A002
        shape = tuple(shape)
        rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        fan_in = shape[-2] * rf if len(shape) >= 2 else shape[0]
        fan_out = shape[-1] * rf if len(shape) >= 2 else shape[0]
        limit = _pymath.sqrt(6.0 / (fan_in + fan_out))
        return _uniform(shape, -limit, limit, dtype)


_INIT_NAMES = {"zeros": Zeros, "ones": Ones, "glorot_uniform": GlorotUniform, "random_normal": RandomNormal,
               "truncated_normal": TruncatedNormal}


def _get_initializer(identifier):
    if identifier is None:
        return None
    if isinstance(identifier, str):
        return _INIT_NAMES[identifier]()
    if isinstance(identifier, type):
        return identifier()
    if callable(identifier):
        return identifier
    raise ValueError(identifier)


zeros_initializer, ones_initializer = Zeros, Ones
initializers.Zeros, initializers.Ones, initializers.Constant = Zeros, Ones, Constant


# ----------------------------------------------------------------------------------------------- activations
def _activation(name):
    if callable(name):
        return name
    table = {
        None: lambda x: x, "linear": lambda x: x,
        "relu": lambda x: torch.relu(x),
        "gelu": lambda x: 0.5 * x * (1.0 + torch.erf(x / _pymath.sqrt(2.0))),  # Keras default: approximate=False
        "swish": lambda x: x * torch.sigmoid(x), "silu": lambda x: x * torch.sigmoid(x),
        "sigmoid": lambda x: torch.sigmoid(x), "tanh": lambda x: torch.tanh(x),
        "softmax": lambda x: torch.softmax(x, -1),
    }
    return table[name]


# ----------------------------------------------------------------------------------------------- Layer
def _snake(name):
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


def _unique(base):
    n = _UIDS.get(base, 0)
    _UIDS[base] = n + 1
    return base if n == 0 else f"{base}_{n}"


def _shape_of(x):
    if isinstance(x, torch.Tensor):
        return TensorShape(x.shape)
    if isinstance(x, np.ndarray):
        return TensorShape(x.shape)
    if isinstance(x, (list, tuple)):
        return [_shape_of(e) for e in x]
    return None


def _cast_inputs(x):
    """Keras autocast: floating inputs are cast to the layer's compute dtype (floatx)."""
    if isinstance(x, np.ndarray):
        x = _t(x)
    if isinstance(x, torch.Tensor):
        x = x if isinstance(x, Tensor) else x.as_subclass(Tensor)
        if x.is_floating_point() and not isinstance(x, Variable) and x.dtype != _dtype(None):
            x = x.to(_dtype(None))
        return x
    if isinstance(x, list):
        return [_cast_inputs(e) for e in x]
    if isinstance(x, tuple):
        return tuple(_cast_inputs(e) for e in x)
    return x


class Layer:
    def __init__(self, trainable=True, name=None, dtype=None, **kwargs):
        if kwargs:
            raise TypeError(f"{type(self).__name__}: unexpected keyword arguments {sorted(kwargs)}")
        self._name = name if name is not None else _unique(_snake(type(self).__name__))
        self.trainable = trainable
        self.built = False
        self._own = []
        self._call_params = None

    # -- attribute tracking: sub-layers, (mutable) lists / dicts of sub-layers and variables, in attribute order,
    #    discovered at access time (Keras wraps list attributes so that later ``.append`` / ``.extend`` are tracked)
    def _tracked_items(self):
        out = []
        for key, value in self.__dict__.items():
            if key.startswith("_own"):
                continue
            if isinstance(value, (Layer, Variable)):
                out.append((key, value))
            elif isinstance(value, (list, tuple, dict)) and len(value):
                vals = list(value.values()) if isinstance(value, dict) else list(value)
                if all(isinstance(v, Layer) for v in vals):
                    out.append((key, value))
        return out

    @property
    def name(self):
        return self._name

    @property
    def dtype(self):
        return _FLOATX[0]

    def _sublayers(self):
        out = []
        for _, v in self._tracked_items():
            if isinstance(v, Layer):
                out.append(v)
            elif isinstance(v, dict):
                out.extend(v.values())
            elif isinstance(v, (list, tuple)):
                out.extend(v)
        return out

    @property
    def layers(self):
        return self._sublayers()

    @property
    def weights(self):
        seen, out = set(), []

        def visit(layer):
            if id(layer) in seen:
                return
            seen.add(id(layer))
            for w in layer._own:
                out.append(w)
            for _, v in layer._tracked_items():
                if isinstance(v, Variable) and not any(v is w for w in out):
                    out.append(v)
            for sub in layer._sublayers():
                visit(sub)

        visit(self)
        return [w for w in out if w.trainable] + [w for w in out if not w.trainable]

    variables = weights

    @property
    def trainable_weights(self):
        return [w for w in self.weights if w.trainable]

    @property
    def non_trainable_weights(self):
        return [w for w in self.weights if not w.trainable]

    def count_params(self):
        return int(sum(w.numel() for w in self.weights))

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, **kwargs):
        init = _get_initializer(initializer) or GlorotUniform()
        dt = _dtype(dtype)
        value = init(tuple(int(s) for s in shape), dtype=dt)
        var = Variable(value, trainable=trainable, name=name, dtype=dt)
        self._own.append(var)
        return var

    def build(self, input_shape):
        self.built = True

    def call(self, inputs, *args, **kwargs):
        return inputs

    def __call__(self, *args, **kwargs):
        if self._call_params is None:
            self._call_params = inspect.signature(self.call).parameters
        params = self._call_params
        if "training" in kwargs and "training" not in params and not any(
                p.kind == p.VAR_KEYWORD for p in params.values()):
            kwargs.pop("training")
        args = tuple(_cast_inputs(a) for a in args)
        with name_scope(self._name), torch.no_grad():
            if not self.built:
                self.build(_shape_of(args[0]) if args else None)
                self.built = True
            return self.call(*args, **kwargs)


# ----------------------------------------------------------------------------------------------- concrete layers
class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.fn = _activation(activation)

    def call(self, x):
        return self.fn(x)


class ReLU(Layer):
    def __init__(self, max_value=None, negative_slope=0.0, threshold=0.0, **kwargs):
        super().__init__(**kwargs)
        assert negative_slope == 0.0 and threshold == 0.0
        self.max_value = max_value

    def call(self, x):
        x = torch.relu(x)
        return x if self.max_value is None else torch.clamp(x, max=float(self.max_value))


class Dropout(Layer):
    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate

    def call(self, x, training=False):
        if training and self.rate > 0:
            raise NotImplementedError("the shim runs inference only")
        return x


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", **kwargs):
        super().__init__(**kwargs)
        self.units, self.use_bias = int(units), use_bias
        self.activation = _activation(activation)
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (input_shape[-1], self.units), initializer=self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.units,), initializer=self.bias_initializer) if self.use_bias else None

    def call(self, x):
        y = x @ self.kernel  # rank > 2: contraction of the last axis (tensordot)
        if self.bias is not None:
            y = y + self.bias
        return self.activation(y)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", data_format=None, dilation_rate=(1, 1),
                 groups=1, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self.filters, self.kernel_size, self.strides = int(filters), _two(kernel_size), _two(strides)
        self.padding, self.dilation_rate, self.groups = padding, _two(dilation_rate), int(groups)
        self.activation, self.use_bias = _activation(activation), use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):
        cin = input_shape[-1]
        self.kernel = self.add_weight("kernel", (*self.kernel_size, cin // self.groups, self.filters),
                                      initializer=self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.filters,), initializer=self.bias_initializer) if self.use_bias else None

    def call(self, x):
        y = _conv2d_nhwc(x, self.kernel, self.strides, self.padding, self.dilation_rate, self.groups)
        if self.bias is not None:
            y = y + self.bias
        return self.activation(y)


class DepthwiseConv2D(Layer):
    def __init__(self, kernel_size, strides=(1, 1), padding="valid", depth_multiplier=1, data_format=None,
                 dilation_rate=(1, 1), groups=1, activation=None, use_bias=True,
                 depthwise_initializer="glorot_uniform", bias_initializer="zeros", depthwise_regularizer=None,
                 bias_regularizer=None, activity_regularizer=None, depthwise_constraint=None, bias_constraint=None,
                 **kwargs):
        super().__init__(**kwargs)
        assert depth_multiplier == 1
        self.kernel_size, self.strides, self.padding = _two(kernel_size), _two(strides), padding
        self.dilation_rate, self.activation, self.use_bias = _two(dilation_rate), _activation(activation), use_bias
        self.depthwise_initializer, self.bias_initializer = depthwise_initializer, bias_initializer

    def build(self, input_shape):
        c = input_shape[-1]
        self.depthwise_kernel = self.add_weight("depthwise_kernel", (*self.kernel_size, c, 1),
                                                initializer=self.depthwise_initializer)
        self.bias = self.add_weight("bias", (c,), initializer=self.bias_initializer) if self.use_bias else None

    def call(self, x):
        c = self.depthwise_kernel.shape[2]
        w = self.depthwise_kernel.reshape(*self.kernel_size, 1, c)
        y = _conv2d_nhwc(x, w, self.strides, self.padding, self.dilation_rate, groups=c)
        if self.bias is not None:
            y = y + self.bias
        return self.activation(y)


class Conv1D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True,
                 kernel_initializer="glorot_uniform", bias_initializer="zeros", **kwargs):
        super().__init__(**kwargs)
        self.filters, self.kernel_size, self.strides = int(filters), int(kernel_size), int(strides)
        self.padding, self.use_bias = padding, use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (self.kernel_size, input_shape[-1], self.filters),
                                      initializer=self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.filters,), initializer=self.bias_initializer) if self.use_bias else None

    def call(self, x):  # (N, L, C)
        y = _conv2d_nhwc(x.unsqueeze(1), self.kernel.unsqueeze(0), (1, self.strides), self.padding).squeeze(1)
        return y + self.bias if self.bias is not None else y


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), **kwargs):
        super().__init__(**kwargs)
        if isinstance(padding, (int, np.integer)):
            self.padding = ((padding, padding), (padding, padding))
        else:
            self.padding = tuple((p, p) if isinstance(p, (int, np.integer)) else tuple(p) for p in padding)

    def call(self, x):
        (pt, pb), (pl, pr) = self.padding
        return F.pad(x, (0, 0, int(pl), int(pr), int(pt), int(pb)))


class ZeroPadding1D(Layer):
    def __init__(self, padding=1, **kwargs):
        super().__init__(**kwargs)
        self.padding = (padding, padding) if isinstance(padding, (int, np.integer)) else tuple(padding)

    def call(self, x):  # (N, L, C)
        return F.pad(x, (0, 0, int(self.padding[0]), int(self.padding[1])))


class LayerNormalization(Layer):
    """Keras' non-fused path (taken for epsilon < 1.001e-5, i.e. both epsilons the reference uses):
    tf.nn.moments over ``axis`` + tf.nn.batch_normalization -- biased variance, eps inside the rsqrt."""

    def __init__(self, axis=-1, epsilon=1e-3, center=True, scale=True, beta_initializer="zeros",
                 gamma_initializer="ones", **kwargs):
        super().__init__(**kwargs)
        assert axis == -1
        self.epsilon, self.center, self.scale = epsilon, center, scale
        self.beta_initializer, self.gamma_initializer = beta_initializer, gamma_initializer

    def build(self, input_shape):
        c = input_shape[-1]
        self.gamma = self.add_weight("gamma", (c,), initializer=self.gamma_initializer) if self.scale else None
        self.beta = self.add_weight("beta", (c,), initializer=self.beta_initializer) if self.center else None

    def call(self, x):
        mean, var = _moments(x, [-1], keepdims=True)
        return _batch_normalization(x, mean, var, self.beta, self.gamma, self.epsilon)


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, beta_initializer="zeros",
                 gamma_initializer="ones", moving_mean_initializer="zeros", moving_variance_initializer="ones",
                 **kwargs):
        super().__init__(**kwargs)
        assert axis == -1
        self.momentum, self.epsilon, self.center, self.scale = momentum, epsilon, center, scale
        self.inits = (gamma_initializer, beta_initializer, moving_mean_initializer, moving_variance_initializer)

    def build(self, input_shape):
        c = input_shape[-1]
        g, b, m, v = self.inits
        self.gamma = self.add_weight("gamma", (c,), initializer=g) if self.scale else None
        self.beta = self.add_weight("beta", (c,), initializer=b) if self.center else None
        self.moving_mean = self.add_weight("moving_mean", (c,), initializer=m, trainable=False)
        self.moving_variance = self.add_weight("moving_variance", (c,), initializer=v, trainable=False)

    def call(self, x, training=False):
        if training:
            raise NotImplementedError("the shim runs inference only")
        return _batch_normalization(x, self.moving_mean, self.moving_variance, self.beta, self.gamma, self.epsilon)


class GlobalAveragePooling2D(Layer):
    def __init__(self, keepdims=False, **kwargs):
        super().__init__(**kwargs)
        self.keepdims = keepdims

    def call(self, x):
        return x.mean(dim=(1, 2), keepdim=self.keepdims)


class GlobalMaxPool2D(Layer):
    def call(self, x):
        return x.amax(dim=(1, 2))


GlobalMaxPooling2D = GlobalMaxPool2D


class GlobalAveragePooling1D(Layer):
    def call(self, x):
        return x.mean(dim=1)


class _Pool2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", **kwargs):
        super().__init__(**kwargs)
        self.pool_size = _two(pool_size)
        self.strides = _two(strides) if strides is not None else self.pool_size
        self.padding = padding

    def _windows(self, x, fill):
        kh, kw = self.pool_size
        sh, sw = self.strides
        if self.padding.lower() == "same":
            pt, pb = _same_pads(x.shape[1], kh, sh)
            pl, pr = _same_pads(x.shape[2], kw, sw)
            x = F.pad(x, (0, 0, pl, pr, pt, pb), value=fill)
        return x.permute(0, 3, 1, 2)


class MaxPool2D(_Pool2D):
    def call(self, x):
        y = F.max_pool2d(self._windows(x, float("-inf")), self.pool_size, self.strides)
        return y.permute(0, 2, 3, 1)


MaxPooling2D = MaxPool2D


class AveragePooling2D(_Pool2D):
    def call(self, x):
        # "same": padded positions are excluded from the average (TF AvgPool divides by the valid count)
        num = F.avg_pool2d(self._windows(x, 0.0), self.pool_size, self.strides, divisor_override=1)
        cnt = F.avg_pool2d(self._windows(torch.ones_like(x[..., :1]), 0.0), self.pool_size, self.strides,
                           divisor_override=1)
        return (num / cnt).permute(0, 2, 3, 1)


class Flatten(Layer):
    def call(self, x):
        return x.reshape(x.shape[0], -1)


class Model(Layer):
    def __init__(self, *args, name=None, **kwargs):
        super().__init__(name=name, **kwargs)

    def save(self, *a, **k):
        raise NotImplementedError("the shim does not serialise models")

    def summary(self, *a, **k):
        print(f"Model {self.name}: {self.count_params()} parameters")


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self._seq = list(layers or [])

    def add(self, layer):
        self._seq = self._seq + [layer]

    def call(self, x, training=False):
        with _fresh_graph():
            for layer in self._seq:
                x = layer(x, training=training)
        return x


# ----------------------------------------------------------------------------------------------- module tree
def _module(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


class _Unsupported:
    """Placeholder for TF symbols outside the hot path: importable, subclassable, unusable."""

    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is not implemented by the TensorFlow shim")

    def __init_subclass__(cls, **kwargs):
        pass


def _lenient(mod):
    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Unsupported,), {})

    mod.__getattr__ = __getattr__
    return mod


_layer_classes = {k: v for k, v in list(globals().items()) if isinstance(v, type) and issubclass(v, Layer)}
layers = _lenient(_module("tensorflow.keras.layers", **_layer_classes))
_keras_init = _lenient(_module(
    "tensorflow.keras.initializers", Initializer=Initializer, Zeros=Zeros, Ones=Ones, Constant=Constant,
    constant=Constant, RandomNormal=RandomNormal, TruncatedNormal=TruncatedNormal, GlorotUniform=GlorotUniform,
    get=_get_initializer))


def _floatx():
    return _FLOATX[0]


def _set_floatx(v):
    assert v in ("float32", "float64")
    _FLOATX[0] = v


def _batch_set_value(tuples):
    for var, value in tuples:
        var.assign(value)


_backend = _module("tensorflow.keras.backend", floatx=_floatx, set_floatx=_set_floatx,
                   batch_set_value=_batch_set_value, clear_session=lambda: _UIDS.clear())


def _register_keras_serializable(package="Custom", name=None):
    return lambda cls: cls


def _load_model(*a, **k):
    raise NotImplementedError("the shim cannot load SavedModels")


_utils = _lenient(_module("tensorflow.keras.utils", register_keras_serializable=_register_keras_serializable))
_models = _lenient(_module("tensorflow.keras.models", load_model=_load_model))
_activations = _module("tensorflow.keras.activations", get=_activation)
_sched = _lenient(_module("tensorflow.keras.optimizers.schedules"))
_optim = _lenient(_module("tensorflow.keras.optimizers", schedules=_sched))
keras = _lenient(_module("tensorflow.keras", layers=layers, initializers=_keras_init, backend=_backend, utils=_utils,
                         models=_models, Model=Model, Sequential=Sequential, activations=_activations,
                         optimizers=_optim))
_pk = _module("tensorflow.python.keras", backend=_backend)
_pf = _lenient(_module("tensorflow.python.framework"))
_module("tensorflow.python.framework.convert_to_constants", convert_variables_to_constants_v2=None)
python = _module("tensorflow.python", keras=_pk, framework=_pf)
for _name, _mod in (("nn", nn), ("linalg", linalg), ("math", math_), ("image", image), ("random", random),
                    ("initializers", initializers)):
    sys.modules[f"tensorflow.{_name}"] = _mod


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return type(name, (_Unsupported,), {})
