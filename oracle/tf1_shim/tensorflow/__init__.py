"""A minimal eager stand-in for the TensorFlow 1.x API surface that hrbigelow/lb-wavenet uses.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (it lives under oracle/; only tests/golden/make_reference_vectors.py
puts it on sys.path).  TensorFlow 1.x cannot be installed in this image (Python 3.12, no network), so the reference's
own Python -- tmodel.py, arch.py, ops.py, ckpt.py, data.py, imported UNMODIFIED from /root/reference -- is executed
against this module instead: every `tf.*` call the reference makes is carried out immediately on torch CPU tensors
(graph mode collapses into program order, which is the order the reference's control dependencies ask for:
tmodel.py:164-166 assign-then-gate, tmodel.py:275-280 print-then-increment).  What this pins is the reference's GRAPH
CONSTRUCTION -- variable names and shapes, concat / slice / dilation / mask / normalisation logic, the order of
operations -- not TensorFlow's kernels: each op below restates the documented TF 1.x semantics in a line or two of
torch (cited per function), in float64 by default so that comparisons with the fp64 oracle are tight.

Only what the reference touches is implemented; anything else raises AttributeError, loudly.
"""
from __future__ import annotations

import builtins
import contextlib
import math
import sys
import types

import numpy as np
import torch

# ---- dtypes ------------------------------------------------------------------------------------------------
_FLOAT = torch.float64  # what tf.float32 maps to (float64: tight oracle comparisons; float32 for the mu-law vectors)
_EAGER = False


def _set_float(dt):
    global _FLOAT, float32
    _FLOAT = dt
    float32 = dt


def _set_eager(flag: bool):
    global _EAGER
    _EAGER = builtins.bool(flag)


float32 = _FLOAT
int32 = torch.int32
int64 = torch.int64
bool = torch.bool  # noqa: A001  (tf.bool)
string = "string"


def executing_eagerly():
    return _EAGER


class TensorShape(list):
    def __init__(self, dims=()):
        super().__init__(dims if isinstance(dims, (list, tuple)) else [dims])


# ---- variables ---------------------------------------------------------------------------------------------
class Variable:
    """tf.Variable / the result of tf.get_variable.  Identity equality (the reference tests `var in vars.values()`,
    arch.py:147); arithmetic reads the current value."""

    def __init__(self, initial_value, dtype=None, name=None, trainable=True):
        v = torch.as_tensor(initial_value)
        if dtype is not None:
            v = v.to(dtype)
        self.name = name or "Variable"
        self.trainable = builtins.bool(trainable)
        self._set(v)

    def _set(self, v):
        v = v.detach().clone()
        if self.trainable and v.dtype.is_floating_point:
            v.requires_grad_(True)
        self._v = v

    def load(self, value):  # tf.Variable.load(value, session)
        self._set(torch.as_tensor(np.asarray(value)).to(self._v.dtype).reshape(self._v.shape))

    @property
    def dtype(self):
        return self._v.dtype

    @property
    def shape(self):
        return tuple(self._v.shape)

    def numpy(self):
        return self._v.detach().numpy()

    def __hash__(self):
        return id(self)

    def __getitem__(self, idx):
        return self._v[idx]

    def __add__(self, o): return self._v + _t(o)
    def __radd__(self, o): return _t(o) + self._v
    def __sub__(self, o): return self._v - _t(o)
    def __rsub__(self, o): return _t(o) - self._v
    def __mul__(self, o): return self._v * _t(o)
    def __rmul__(self, o): return _t(o) * self._v
    def __truediv__(self, o): return self._v / _t(o)
    def __mod__(self, o): return self._v % _t(o)
    def __neg__(self): return -self._v


def _t(x):
    """anything the reference passes as a 'tensor' -> torch tensor (python scalars stay scalars)"""
    if isinstance(x, Variable):
        return x._v
    if isinstance(x, (int, float, np.integer, np.floating, builtins.bool)):
        return x
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return t.to(_FLOAT) if t.dtype.is_floating_point else t
    return x


_var_store: dict = {}
_scope: list = []


def _reset():
    """forget every variable (a fresh default graph)"""
    _var_store.clear()
    del _scope[:]


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


@contextlib.contextmanager
def name_scope(name, *a, **k):  # does not prefix tf.get_variable names in TF 1.x
    yield


@contextlib.contextmanager
def control_dependencies(ops):  # eager: the listed ops have already run
    yield


class zeros_initializer:
    def __init__(self, dtype=None):
        pass

    def __call__(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)


class constant_initializer:
    def __init__(self, value=0.0, dtype=None):
        self.value = value

    def __call__(self, shape, dtype):
        return torch.full(shape, self.value, dtype=dtype)


_init_gen = torch.Generator().manual_seed(0)


class _XavierConv2d:
    """tf.contrib.layers.xavier_initializer_conv2d(uniform=True) = variance_scaling(1.0, FAN_AVG, uniform):
    U(-l, l), l = sqrt(6 / (fan_in + fan_out)), fan_in = shape[-2] * prod(shape[:-2]), fan_out = shape[-1] * prod(shape[:-2])"""

    def __call__(self, shape, dtype):
        shape = list(shape)
        rec = 1
        for s in shape[:-2]:
            rec *= s
        fan_in = (shape[-2] if len(shape) > 1 else shape[-1]) * rec
        fan_out = shape[-1] * rec
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=_init_gen, dtype=torch.float64) * 2 - 1).to(dtype) * lim


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
    """tf.get_variable under the current variable_scope stack; an existing name is returned as is (AUTO_REUSE), which
    is how a second call of WaveNetTrain.build() runs the next stage on the same variables."""
    full = "/".join(_scope + [name])
    if full in _var_store:
        return _var_store[full]
    dtype = dtype or float32
    if isinstance(shape, (int, np.integer)):
        shape = [int(shape)]
    shape = [int(s) for s in (shape or [])]
    if initializer is None:
        initializer = _XavierConv2d()
    if isinstance(initializer, type):  # a class, e.g. tf.zeros_initializer (tmodel.py:224): TF instantiates it
        initializer = initializer()
    val = initializer(shape, dtype) if callable(initializer) else torch.as_tensor(initializer, dtype=dtype)
    v = Variable(val, dtype=dtype, name=full + ":0", trainable=trainable)
    _var_store[full] = v
    return v


def assign(ref, value, *a, **k):
    ref._set(torch.as_tensor(_t(value)).to(ref._v.dtype))
    return ref._v


def variables_initializer(var_list, name=None):
    return None


# ---- ops (TF 1.x python API semantics) ------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    if isinstance(value, (int, np.integer)) and dtype is None:
        return torch.tensor(int(value), dtype=torch.int32)
    t = torch.as_tensor(value)
    if dtype is not None:
        t = t.to(dtype)
    elif t.dtype.is_floating_point:
        t = t.to(_FLOAT)
    return t


def identity(x, name=None):
    return _t(x)


def stop_gradient(x, name=None):
    return _t(x).detach()


def cast(x, dtype, name=None):
    x = _t(x)
    return x.to(dtype) if torch.is_tensor(x) else torch.tensor(x, dtype=dtype)


def to_float(x, name=None):
    return cast(x, float32)


def to_int32(x, name=None):  # float -> int casts truncate toward zero
    return cast(x, torch.int32)


def shape(x, name=None):
    return torch.tensor(list(_t(x).shape), dtype=torch.int32)


def add(x, y, name=None):
    return _t(x) + _t(y)


def add_n(inputs, name=None):
    out = _t(inputs[0])
    for x in inputs[1:]:
        out = out + _t(x)
    return out


def abs(x, name=None):  # noqa: A001
    return torch.abs(_t(x))


def sign(x, name=None):
    return torch.sign(_t(x))


def log1p(x, name=None):
    x = _t(x)
    return torch.log1p(x if torch.is_tensor(x) else torch.tensor(x, dtype=_FLOAT))


def tanh(x, name=None):
    return torch.tanh(_t(x))


def sigmoid(x, name=None):
    return torch.sigmoid(_t(x))


def equal(x, y, name=None):
    return torch.as_tensor(_t(x) == _t(y))


def not_equal(x, y, name=None):
    return torch.as_tensor(_t(x) != _t(y))


def less(x, y, name=None):
    return torch.as_tensor(_t(x) < _t(y))


def reduce_sum(x, axis=None, name=None):
    x = _t(x)
    return x.sum() if axis is None else x.sum(axis)


def reduce_mean(x, axis=None, name=None):
    """integer inputs: TF's Mean kernel divides in the input type (truncating), tmodel.py:244"""
    x = _t(x)
    if not x.dtype.is_floating_point:
        s = x.sum() if axis is None else x.sum(axis)
        n = x.numel() if axis is None else x.shape[axis]
        return torch.div(s, n, rounding_mode="trunc").to(x.dtype)
    return x.mean() if axis is None else x.mean(axis)


def argmax(x, axis=None, output_type=torch.int64, name=None):
    return torch.argmax(_t(x), dim=axis).to(output_type)


def concat(values, axis, name=None):
    return torch.cat([_t(v) for v in values], dim=axis)


def stack(values, axis=0, name=None):
    return torch.stack([torch.as_tensor(_t(v)) for v in values], dim=axis)


def expand_dims(x, axis, name=None):
    return _t(x).unsqueeze(axis)


def squeeze(x, axis=None, name=None):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def broadcast_to(x, shape_, name=None):
    return _t(x).expand(*[int(s) for s in shape_])


def zeros(shape_, dtype=None, name=None):
    return torch.zeros([int(s) for s in shape_], dtype=dtype or float32)


def gather(params, indices, name=None):  # axis 0
    return _t(params)[_t(indices).long()]


def matmul(a, b, name=None):  # batched over leading dimensions
    return torch.matmul(_t(a), _t(b))


def one_hot(indices, depth, axis=-1, name=None, dtype=None):
    """indices outside [0, depth) give an all-zero row (tf.one_hot documentation)"""
    idx = _t(indices).long()
    out = (idx.unsqueeze(-1) == torch.arange(int(depth))).to(dtype or float32)
    assert axis == -1
    return out


def cond(pred, true_fn=None, false_fn=None, name=None):
    p = _t(pred)
    return true_fn() if builtins.bool(p) else false_fn()


def py_func(func, inp, Tout, name=None):
    args = []
    for x in inp:
        x = _t(x)
        args.append(x.detach().numpy()[()] if torch.is_tensor(x) else x)
    return func(*args)


def Print(input_, data, message=None, **k):
    return _t(input_)


def random_uniform(shape_, minval=0, maxval=1, dtype=None, **k):
    return torch.rand([int(s) for s in shape_], generator=_init_gen, dtype=torch.float64).to(dtype or float32) * (maxval - minval) + minval


class GradientTape:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, target, sources):
        return train.Optimizer._grads(target, sources)


# ---- tf.nn ------------------------------------------------------------------------------------------------------
def _nn_convolution(input, filter, padding, strides=None, dilation_rate=None, name=None, data_format=None):  # noqa: A002
    """1-D tf.nn.convolution on NWC input with a [width, in, out] filter: a cross-correlation,
    out[b, t, o] = sum_{k, c} input[b, t + k * dilation, c] * filter[k, c, o], 'VALID' = no padding"""
    assert padding == "VALID" and (strides is None or list(strides) == [1])
    x, w = _t(input), _t(filter)
    dil = int(dilation_rate[0]) if dilation_rate is not None else 1
    y = torch.nn.functional.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), dilation=dil)
    return y.transpose(1, 2)


class _XentWithLogits(torch.autograd.Function):
    """TF's SoftmaxCrossEntropyWithLogits kernel (core/kernels/xent_op.h) and its registered gradient
    (python/ops/nn_grad.py): the op returns (loss, backprop) with
        loss = sum_q labels * (log sum exp(logits - max) - (logits - max)),   backprop = softmax - labels,
    and d loss / d logits = grad_loss * backprop WHATEVER the labels sum to -- for the all-zero row that tf.one_hot makes
    of an out-of-range code this is softmax, not the zero that differentiating the formula would give;
    d loss / d labels = grad_loss * (-log_softmax)."""

    @staticmethod
    def forward(ctx, logits, labels):
        lsm = torch.log_softmax(logits, dim=-1)
        ctx.save_for_backward(torch.exp(lsm) - labels, lsm)
        return -(labels * lsm).sum(-1)

    @staticmethod
    def backward(ctx, g):
        backprop, lsm = ctx.saved_tensors
        return g.unsqueeze(-1) * backprop, g.unsqueeze(-1) * (-lsm)


def _nn_softmax_xent_v2(labels=None, logits=None, dim=-1, name=None, axis=None):
    d = dim if axis is None else axis
    lg, lb = _t(logits), _t(labels)
    assert d in (-1, lg.dim() - 1)
    return _XentWithLogits.apply(lg, lb)


def _nn_conv1d_transpose(value, filter, output_shape, stride, padding="SAME", data_format="NWC", name=None):  # noqa: A002
    """tf.contrib.nn.conv1d_transpose: value [B, T, in], filter [width, out, in], 'SAME': output length T * stride,
    out[b, t * stride + k - pad_left, o] += value[b, t, c] * filter[k, o, c], pad_left = max(width - stride, 0) // 2"""
    x, w = _t(value), _t(filter)
    s = int(stride)
    width = w.shape[0]
    y = torch.nn.functional.conv_transpose1d(x.transpose(1, 2), w.permute(2, 1, 0), stride=s)  # [B, out, (T-1)s + width]
    want = int(x.shape[1]) * s
    pad_left = max(width - s, 0) // 2
    y = y[:, :, pad_left:pad_left + want]
    out_shape = [int(v) for v in _t(output_shape)]
    y = y.transpose(1, 2)
    assert list(y.shape) == out_shape, (list(y.shape), out_shape)
    return y


nn = types.SimpleNamespace(
    convolution=_nn_convolution,
    relu=lambda x, name=None: torch.relu(_t(x)),
    softmax=lambda x, axis=-1, name=None, dim=None: torch.softmax(_t(x), dim=axis if dim is None else dim),
    l2_loss=lambda x, name=None: (_t(x) ** 2).sum() / 2,  # sum(t ** 2) / 2
    softmax_cross_entropy_with_logits_v2=_nn_softmax_xent_v2,
)


def multinomial(logits, num_samples, **k):
    raise NotImplementedError("tf.multinomial: TF's sampler is not reproducible outside TF")


# ---- tf.train / tf.summary / tf.contrib / tf.errors / tf.data ------------------------------------------------
class _Optimizer:
    def __init__(self, use_locking, name):
        pass

    @staticmethod
    def _grads(loss, var_list):
        gs = torch.autograd.grad(_t(loss), [v._v for v in var_list], allow_unused=True, retain_graph=True)
        return list(gs)

    def compute_gradients(self, loss, var_list=None, **k):
        """[(d loss / d var, var)] -- tf.gradients on the scalar loss (None for unreachable variables)"""
        return list(zip(self._grads(loss, var_list), var_list))


class _Saver:
    def __init__(self, *a, **k):
        pass


train = types.SimpleNamespace(Optimizer=_Optimizer, Saver=_Saver)
summary = types.SimpleNamespace(histogram=lambda *a, **k: None, scalar=lambda *a, **k: None)


class _OutOfRangeError(Exception):
    pass


errors = types.SimpleNamespace(OutOfRangeError=_OutOfRangeError)

contrib = types.ModuleType("tensorflow.contrib")
contrib.nn = types.SimpleNamespace(conv1d_transpose=_nn_conv1d_transpose)
contrib.layers = types.SimpleNamespace(xavier_initializer_conv2d=lambda *a, **k: _XavierConv2d())
if __name__ == "tensorflow":  # `import tensorflow.contrib.eager` style imports of the reference
    sys.modules["tensorflow.contrib"] = contrib
