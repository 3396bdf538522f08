"""TEST INFRASTRUCTURE ONLY -- a torch-backed stand-in for the slice of the
`tensorflow` Python API that thiagopbueno/tf-mpc v0.7.0 touches.

Why it exists: TensorFlow is not installable in the build container (no wheel,
no network), so the reference cannot be imported as shipped.  With this package
first on ``sys.path`` the UNMODIFIED reference sources under ``/root/reference``
import and run: every line of the reference's own Python control flow
(``tfmpc/solvers/lqr.py``, ``ilqr.py``, ``utils/optimization.py``, the envs and
``DiffEnv``'s GradientTape calls) executes verbatim, with TensorFlow's CPU
kernels replaced by torch's CPU kernels of the same mathematical definition.
``tests/golden/make_golden.py`` uses it to generate the golden fixtures that pin
the oracle; nothing in the product package imports it.

What it pins: the reference's algorithm and branch decisions.  What it does NOT
pin: Eigen-vs-torch rounding of individual fp32 ops (summation order inside
matmul / Cholesky), which is below the parity tolerances.

Set ``TF_SHIM_FLOAT=float64`` before import to run the reference in double
precision (every "float32" request maps to float64).
"""
import os as _os
import sys as _sys
import types as _types

import numpy as _np
import torch as _torch

_FLOAT = _torch.float64 if _os.environ.get("TF_SHIM_FLOAT", "float32") == "float64" else _torch.float32

float32 = _FLOAT
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
bool = _torch.bool  # noqa: A001  (mirrors tf.bool)

__version__ = "0.0-torch-shim"


class TensorShape(tuple):
    """tf.TensorShape compares equal to lists and tuples alike."""

    def __eq__(self, other):
        try:
            other = tuple(other)
        except TypeError:
            return False
        if len(other) != len(self):
            return False
        return all(int(a) == int(b) for a, b in zip(self, other))

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = tuple.__hash__

    def as_list(self):
        return list(self)


_UFUNC_TO_OP = {"add": "__add__", "subtract": "__sub__", "multiply": "__mul__", "true_divide": "__truediv__",
                "divide": "__truediv__", "less": "__lt__", "less_equal": "__le__", "greater": "__gt__",
                "greater_equal": "__ge__", "equal": "__eq__", "not_equal": "__ne__", "power": "__pow__"}


def _conv_other(other, like):
    if isinstance(other, _np.ndarray) or isinstance(other, (list, tuple)):
        t = _torch.as_tensor(_np.asarray(other))
        if t.is_floating_point():
            t = t.to(like.dtype if like.is_floating_point() else _FLOAT)
        return t.as_subclass(Tensor)
    if isinstance(other, _np.generic):
        return other.item()
    return other


class Tensor(_torch.Tensor):
    """torch.Tensor with the handful of tf.Tensor behaviours the reference relies on."""

    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        # tf.Tensor outranks ndarray in mixed arithmetic (ndarray <op> Tensor -> Tensor);
        # every other ufunc sees plain arrays, as tf.Tensor.__array__ would give it.
        if method == "__call__" and len(inputs) == 2 and ufunc.__name__ in _UFUNC_TO_OP and not kwargs:
            a, b = inputs
            like = a if isinstance(a, _torch.Tensor) else b
            a = a if isinstance(a, _torch.Tensor) else _conv_other(a, like)
            b = b if isinstance(b, _torch.Tensor) else _conv_other(b, like)
            if not isinstance(a, _torch.Tensor):
                a = _torch.as_tensor(a, dtype=like.dtype).as_subclass(Tensor)
            return getattr(a, _UFUNC_TO_OP[ufunc.__name__])(b)
        inputs = tuple(i.numpy() if isinstance(i, _torch.Tensor) else i for i in inputs)
        return getattr(ufunc, method)(*inputs, **kwargs)

    # tf.Tensor is immutable: `a += b` rebinds the name, it never mutates `a` in place.
    def __iadd__(self, other):
        return self.__add__(other)

    def __isub__(self, other):
        return self.__sub__(other)

    def __imul__(self, other):
        return self.__mul__(other)

    def __itruediv__(self, other):
        return self.__truediv__(other)

    @property
    def shape(self):
        return TensorShape(_torch.Tensor.size(self))

    def numpy(self):
        return _torch.Tensor.numpy(self.detach().as_subclass(_torch.Tensor))

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __iter__(self):
        return (self[i] for i in builtins_range(_torch.Tensor.size(self, 0)))

    def __hash__(self):
        return id(self)

    def sum(self, *args, axis=None, out=None, keepdims=False, **kw):
        # numpy's np.sum(obj) forwards to obj.sum(axis=..., out=...)
        if args or kw:
            return _torch.Tensor.sum(self, *args, **kw)
        if axis is None:
            return _torch.Tensor.sum(self)
        return _torch.Tensor.sum(self, dim=axis, keepdim=builtins_bool(keepdims) if keepdims in (True, False) else False)

    def __format__(self, spec):
        if self.numel() == 1:
            return self.detach().as_subclass(_torch.Tensor).reshape(-1)[0].item().__format__(spec)
        return str(self)

    def __int__(self):
        return int(self.detach().as_subclass(_torch.Tensor).reshape(-1)[0].item())

    def __index__(self):
        return self.__int__()

    def __float__(self):
        return float(self.detach().as_subclass(_torch.Tensor).reshape(-1)[0].item())

    # python truthiness of a 1-element tensor, as TF eager does
    def __bool__(self):
        return builtins_bool(self.detach().as_subclass(_torch.Tensor).reshape(-1)[0].item()) \
            if self.numel() == 1 else _torch.Tensor.__bool__(self)


def _wrap_binary(name):
    base = getattr(_torch.Tensor, name)

    def op(self, other):
        return base(self, _conv_other(other, self))
    op.__name__ = name
    return op


import builtins as _builtins  # noqa: E402
builtins_bool = _builtins.bool

for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__",
           "__truediv__", "__rtruediv__", "__pow__", "__rpow__",
           "__lt__", "__le__", "__gt__", "__ge__", "__eq__", "__ne__",
           "__and__", "__or__"):
    setattr(Tensor, _n, _wrap_binary(_n))


class Variable(Tensor):
    def __new__(cls, initial_value=None, trainable=None, dtype=None, name=None, **kw):
        t = convert_to_tensor(initial_value, dtype=dtype).clone()
        return t.as_subclass(cls)

    def __init__(self, *a, **kw):
        pass

    def assign(self, value):
        with _torch.no_grad():
            self.copy_(convert_to_tensor(value, dtype=self.dtype).reshape(_torch.Tensor.size(self)))
        return self

    def scatter_nd_update(self, indices, updates):
        idx = convert_to_tensor(indices).long()
        upd = convert_to_tensor(updates, dtype=self.dtype)
        if idx.numel() == 0:
            return self
        with _torch.no_grad():
            _torch.Tensor.__setitem__(self, tuple(idx.t()), upd)
        return self


def _t(x):
    return x.as_subclass(Tensor) if isinstance(x, _torch.Tensor) else x


def convert_to_tensor(value, dtype=None, dtype_hint=None):
    if isinstance(value, _torch.Tensor):
        t = value
    elif isinstance(value, _np.ndarray) or isinstance(value, _np.generic):
        t = _torch.as_tensor(_np.array(value))
        if t.is_floating_point():
            t = t.to(_FLOAT)
    else:
        t = _torch.as_tensor(_np.array(_tolist(value)))
        if t.is_floating_point():
            t = t.to(_FLOAT)
        elif t.dtype == _torch.int64:
            t = t.to(_torch.int32)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.as_subclass(Tensor)


def _tolist(v):
    if isinstance(v, _torch.Tensor):
        return v.detach().cpu().tolist()
    if isinstance(v, (list, tuple)):
        return [_tolist(e) for e in v]
    if isinstance(v, (_np.ndarray, _np.generic)):
        return v.tolist()
    return v


def constant(value, dtype=None, shape=None, name=None):
    t = convert_to_tensor(value, dtype=dtype)
    if shape is not None:
        shp = [int(s) for s in _tolist(shape)] if not isinstance(shape, int) else [shape]
        if t.numel() == 1 and int(_np.prod(shp)) != 1:
            t = t.reshape(()).expand(shp).clone()
        else:
            t = t.reshape(shp)
    return t.clone().as_subclass(Tensor) if isinstance(value, _torch.Tensor) else t


def function(fn=None, **kw):
    """tf.function: graph compilation is an execution detail; run eagerly."""
    if fn is None:
        return lambda f: f
    return fn


def py_function(func, inp, Tout=None):
    return func(*inp)


# ---------------------------------------------------------------- elementwise / shape
def _c(x, like=None):
    if isinstance(x, _torch.Tensor):
        return x
    t = convert_to_tensor(x)
    if like is not None and isinstance(like, _torch.Tensor) and t.is_floating_point() and like.is_floating_point():
        t = t.to(like.dtype)
    return t


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _c(a), _c(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return _t(_torch.matmul(a, b))


def transpose(a, perm=None):
    a = _c(a)
    if perm is None:
        perm = list(builtins_range(a.dim()))[::-1]
    return _t(a.permute(*perm))


def _axis(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


def squeeze(x, axis=None):
    x = _c(x)
    if axis is None:
        return _t(x.squeeze())
    ax = _axis(axis)
    if isinstance(ax, int):
        ax = (ax,)
    ax = sorted(a % x.dim() for a in ax)
    for a in ax:
        if x.shape[a] != 1:
            raise errors.InvalidArgumentError(f"cannot squeeze dim {a} of size {x.shape[a]}")
    for a in reversed(ax):
        x = x.squeeze(a)
    return _t(x)


def expand_dims(x, axis):
    return _t(_c(x).unsqueeze(int(axis)))


def reshape(x, shape):
    shape = [int(s) for s in _tolist(shape)]
    return _t(_c(x).reshape(shape))


def shape(x):  # noqa: A001
    return constant(list(_c(x).shape), dtype=int32)


def cast(x, dtype):
    return _t(_c(x).to(dtype))


def concat(values, axis):
    return _t(_torch.cat([_c(v) for v in values], dim=int(axis)))


def stack(values, axis=0):
    return _t(_torch.stack([_c(v) for v in values], dim=int(axis)))


def eye(n, dtype=None):
    return _t(_torch.eye(int(n), dtype=dtype or _FLOAT))


def zeros(shape, dtype=None):
    return _t(_torch.zeros([int(s) for s in _tolist(shape)], dtype=dtype or _FLOAT))


def ones(shape, dtype=None):
    return _t(_torch.ones([int(s) for s in _tolist(shape)], dtype=dtype or _FLOAT))


def zeros_like(x):
    return _t(_torch.zeros_like(_c(x)))


def ones_like(x):
    return _t(_torch.ones_like(_c(x)))


def range(start, limit=None, delta=1):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _t(_torch.arange(int(start), int(limit), int(delta), dtype=_torch.int32))


def where(condition, x=None, y=None):
    condition = _c(condition)
    if x is None and y is None:
        return _t(_torch.nonzero(condition))
    x = _c(x)
    y = _c(y, like=x)
    x = _c(x, like=y)
    return _t(_torch.where(condition, x, y))


def abs(x):  # noqa: A001
    return _t(_torch.abs(_c(x)))


def sign(x):
    return _t(_torch.sign(_c(x)))


def exp(x):
    return _t(_torch.exp(_c(x)))


def sin(x):
    return _t(_torch.sin(_c(x)))


def cos(x):
    return _t(_torch.cos(_c(x)))


def sqrt(x):
    return _t(_torch.sqrt(_c(x)))


def maximum(x, y):
    x = _c(x) if isinstance(x, _torch.Tensor) else x
    y = _c(y) if isinstance(y, _torch.Tensor) else y
    if not isinstance(x, _torch.Tensor):
        x = _torch.as_tensor(x, dtype=y.dtype)
    if not isinstance(y, _torch.Tensor):
        y = _torch.as_tensor(y, dtype=x.dtype)
    # TF's MaximumGrad routes the gradient to x where x >= y, else to y; torch
    # splits ties 50/50.  Use where() so the tie rule is TF's.
    return _t(_torch.where(x >= y, x.expand(_torch.broadcast_shapes(x.shape, y.shape)),
                           y.expand(_torch.broadcast_shapes(x.shape, y.shape))))


def clip_by_value(x, lo, hi):
    x = _c(x)
    lo = _c(lo, like=x)
    hi = _c(hi, like=x)
    if lo.is_floating_point():
        lo = lo.to(x.dtype)
    if hi.is_floating_point():
        hi = hi.to(x.dtype)
    return _t(_torch.minimum(_torch.maximum(x, lo), hi))


def logical_or(a, b):
    return _t(_torch.logical_or(_c(a), _c(b)))


def logical_and(a, b):
    return _t(_torch.logical_and(_c(a), _c(b)))


def logical_not(a):
    return _t(_torch.logical_not(_c(a)))


def _reduce(fn, x, axis, keepdims):
    x = _c(x)
    ax = _axis(axis)
    if ax is None:
        return _t(fn(x))
    return _t(fn(x, dim=ax, keepdim=keepdims))


def reduce_sum(x, axis=None, keepdims=False):
    return _reduce(_torch.sum, x, axis, keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    return _reduce(_torch.mean, x, axis, keepdims)


def reduce_prod(x, axis=None, keepdims=False):
    x = _c(x)
    if axis is None:
        return _t(_torch.prod(x))
    return _t(_torch.prod(x, dim=int(axis), keepdim=keepdims))


def reduce_max(x, axis=None, keepdims=False):
    x = _c(x)
    if axis is None:
        return _t(_torch.max(x))
    return _t(_torch.amax(x, dim=_axis(axis), keepdim=keepdims))


def reduce_all(x, axis=None):
    x = _c(x)
    if x.dtype != _torch.bool:
        x = x != 0
    return _t(_torch.all(x)) if axis is None else _t(_torch.all(x, dim=int(axis)))


def reduce_any(x, axis=None):
    x = _c(x)
    if x.dtype != _torch.bool:
        x = x != 0
    return _t(_torch.any(x)) if axis is None else _t(_torch.any(x, dim=int(axis)))


def norm(x, axis=None):
    x = _c(x)
    if axis is None:
        return _t(_torch.sqrt(_torch.sum(x * x)))
    return _t(_torch.sqrt(_torch.sum(x * x, dim=int(axis))))


def gather(params, indices, axis=0):
    return _t(_torch.index_select(_c(params), int(axis), _c(indices).long().reshape(-1)))


def tensordot(a, b, axes):
    return _t(_torch.tensordot(_c(a), _c(b), dims=axes))


def roll(x, shift, axis):
    return _t(_torch.roll(_c(x), int(shift), int(axis)))


class _Math(_types.ModuleType):
    pass


math = _Math("tensorflow.math")
math.maximum = maximum
math.is_inf = lambda x: _t(_torch.isinf(_c(x)))
math.count_nonzero = lambda x, axis=None: _t(_torch.count_nonzero(_c(x)))
math.abs = abs
math.exp = exp
math.sin = sin
math.cos = cos
_sys.modules["tensorflow.math"] = math


class _InvalidArgumentError(Exception):
    pass


errors = _types.ModuleType("tensorflow.errors")
errors.InvalidArgumentError = _InvalidArgumentError
_sys.modules["tensorflow.errors"] = errors


# ---------------------------------------------------------------- linalg
linalg = _types.ModuleType("tensorflow.linalg")


def _cholesky(a):
    a = _c(a)
    if a.numel() == 0:
        return _t(a.clone())
    L, info = _torch.linalg.cholesky_ex(a)
    if int(info.max()) != 0 or not builtins_bool(_torch.isfinite(L).all()):
        # TF's CPU kernel (Eigen LLT) reports failure as InvalidArgumentError.
        raise errors.InvalidArgumentError("Cholesky decomposition was not successful. The input might not be valid.")
    return _t(L)


def _cholesky_solve(chol, rhs):
    chol, rhs = _c(chol), _c(rhs)
    if chol.numel() == 0:
        return _t(rhs.clone())
    return _t(_torch.cholesky_solve(rhs, chol, upper=False))


def _inv(a):
    a = _c(a)
    try:
        return _t(_torch.linalg.inv(a))
    except Exception as e:  # singular
        raise errors.InvalidArgumentError(str(e))


def _band_part(x, num_lower, num_upper):
    x = _c(x)
    n, m = x.shape[-2], x.shape[-1]
    i = _torch.arange(n).reshape(-1, 1)
    j = _torch.arange(m).reshape(1, -1)
    keep = _torch.ones(n, m, dtype=_torch.bool)
    if num_lower >= 0:
        keep &= (i - j) <= num_lower
    if num_upper >= 0:
        keep &= (j - i) <= num_upper
    return _t(_torch.where(keep, x, _torch.zeros_like(x)))


linalg.cholesky = _cholesky
linalg.cholesky_solve = _cholesky_solve
linalg.inv = _inv
linalg.band_part = _band_part
linalg.diag = lambda d: _t(_torch.diag_embed(_c(d)))
linalg.diag_part = lambda x: _t(_torch.diagonal(_c(x), dim1=-2, dim2=-1))
linalg.matrix_transpose = lambda x: _t(_c(x).transpose(-1, -2))
linalg.norm = norm
_sys.modules["tensorflow.linalg"] = linalg


# ---------------------------------------------------------------- random
random = _types.ModuleType("tensorflow.random")
_gen = _torch.Generator().manual_seed(0)


def _shape_list(shape):
    return [int(s) for s in _tolist(shape)]


def _uniform(shape, minval=0, maxval=None, dtype=None, seed=None):
    dtype = dtype or _FLOAT
    shp = _shape_list(shape)
    if dtype in (int32, int64):
        return _t(_torch.randint(int(minval), int(maxval), shp, generator=_gen, dtype=dtype))
    if maxval is None:
        maxval = 1.0
    r = _torch.rand(shp, generator=_gen, dtype=dtype)
    lo = _c(minval).to(dtype) if not isinstance(minval, (int, float)) else minval
    hi = _c(maxval).to(dtype) if not isinstance(maxval, (int, float)) else maxval
    # TF: rnd * (maxval - minval) + minval, with broadcasting of tensor-valued bounds
    return _t(r * (hi - lo) + lo)


def _normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
    dtype = dtype or _FLOAT
    return _t(_torch.randn(_shape_list(shape), generator=_gen, dtype=dtype) * stddev + mean)


def _truncated_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
    dtype = dtype or _FLOAT
    shp = _shape_list(shape)
    out = _torch.randn(shp, generator=_gen, dtype=dtype)
    bad = out.abs() > 2
    while builtins_bool(bad.any()):
        out = _torch.where(bad, _torch.randn(shp, generator=_gen, dtype=dtype), out)
        bad = out.abs() > 2
    return _t(out * stddev + mean)


def _gamma(shape, alpha, beta=None, dtype=None, seed=None):
    alpha = _c(alpha)
    beta = _c(beta) if beta is not None else _torch.ones_like(alpha)
    d = _torch.distributions.Gamma(alpha, beta)
    return _t(d.sample(_shape_list(shape)))


def set_seed(seed):
    _gen.manual_seed(int(seed))


random.uniform = _uniform
random.normal = _normal
random.truncated_normal = _truncated_normal
random.gamma = _gamma
random.set_seed = set_seed
_sys.modules["tensorflow.random"] = random


# ---------------------------------------------------------------- TensorArray
class TensorArray:
    def __init__(self, dtype=None, size=0, **kw):
        self._items = [None] * int(size)

    def write(self, index, value):
        self._items[int(index)] = _c(value)
        return self

    def read(self, index):
        return _t(self._items[int(index)])

    def stack(self):
        return _t(_torch.stack(self._items, dim=0))


# ---------------------------------------------------------------- autodiff
class UnconnectedGradients:
    NONE = "none"
    ZERO = "zero"


class GradientTape:
    """Reverse-mode tape over torch.autograd.

    ``watch`` flips requires_grad on the very tensor object the reference then
    feeds to its env functions, so the torch graph plays the role of the tape.
    Every gradient is taken with create_graph=True so that the reference's
    second-order calls (Jacobians of l_x, l_u) see a differentiable l_x.
    """

    def __init__(self, persistent=False, watch_accessed_variables=True):
        self._watched = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, tensor):
        if not tensor.requires_grad:
            tensor.requires_grad_(True)
        self._watched.append(tensor)

    @staticmethod
    def _grad(y, xs, zero):
        outs = []
        if not y.requires_grad:
            grads = [None] * len(xs)
        else:
            grads = _torch.autograd.grad(y, xs, retain_graph=True, create_graph=True, allow_unused=True)
        for g, x in zip(grads, xs):
            if g is None:
                g = _torch.zeros_like(x) if zero else None
            outs.append(g)
        return outs

    def gradient(self, target, sources, unconnected_gradients=UnconnectedGradients.NONE):
        single = not isinstance(sources, (list, tuple))
        xs = [sources] if single else list(sources)
        zero = unconnected_gradients == UnconnectedGradients.ZERO
        # TF sums the target when it is not a scalar.
        res = self._grad(_c(target).sum(), xs, zero)
        res = [_t(g) if g is not None else None for g in res]
        return res[0] if single else res

    def jacobian(self, target, sources, unconnected_gradients=UnconnectedGradients.NONE):
        single = not isinstance(sources, (list, tuple))
        xs = [sources] if single else list(sources)
        target = _c(target)
        flat = target.reshape(-1)
        rows = [[] for _ in xs]
        for i in builtins_range(flat.numel()):
            gs = self._grad(flat[i], xs, True)
            for r, g in zip(rows, gs):
                r.append(g)
        out = []
        for r, x in zip(rows, xs):
            if len(r) == 0:
                out.append(_t(_torch.zeros(list(target.shape) + list(x.shape), dtype=x.dtype)))
            else:
                out.append(_t(_torch.stack(r, 0).reshape(list(target.shape) + list(x.shape))))
        return out[0] if single else out

    def batch_jacobian(self, target, source, unconnected_gradients=UnconnectedGradients.NONE):
        target = _c(target)
        B = target.shape[0]
        tshape = list(target.shape[1:])
        sshape = list(source.shape[1:])
        flat = target.reshape(B, -1)
        cols = []
        for i in builtins_range(flat.shape[1]):
            # rows of the batch are independent, so d(sum_b y[b,i])/dx[b] = dy[b,i]/dx[b]
            (g,) = self._grad(flat[:, i].sum(), [source], True)
            cols.append(g.reshape(B, -1))
        J = _torch.stack(cols, dim=1) if cols else _torch.zeros(B, 0, int(_np.prod(sshape)))
        return _t(J.reshape([B] + tshape + sshape))


builtins_range = _builtins.range

# ---------------------------------------------------------------- compat.v1.logging
import logging as _logging  # noqa: E402

compat = _types.ModuleType("tensorflow.compat")
compat.v1 = _types.ModuleType("tensorflow.compat.v1")
_tfl = _types.ModuleType("tensorflow.compat.v1.logging")
_logger = _logging.getLogger("tensorflow")
_tfl.DEBUG, _tfl.INFO, _tfl.WARN, _tfl.ERROR, _tfl.FATAL = 10, 20, 30, 40, 50
_tfl.set_verbosity = lambda v: _logger.setLevel(v)
# The reference builds f-strings with .numpy().tolist() for every debug line; the
# arguments are evaluated by the caller regardless, we only drop the message.
_tfl.debug = lambda *a, **k: None
_tfl.info = lambda *a, **k: None
_tfl.warn = lambda *a, **k: None
_tfl.warning = lambda *a, **k: None
_tfl.error = lambda msg, *a, **k: _logger.error(msg)
compat.v1.logging = _tfl
_sys.modules["tensorflow.compat"] = compat
_sys.modules["tensorflow.compat.v1"] = compat.v1
_sys.modules["tensorflow.compat.v1.logging"] = _tfl
