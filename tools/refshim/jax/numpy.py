"""jax.numpy on numpy: float32 / int32 defaults, `.at[...]` functional updates."""
import numpy as _np
from numpy import pi, inf, nan, newaxis  # noqa: F401


class _ScalarType:
    """jnp.float32 & co: callable scalar constructors that also carry `.dtype` (control_environment_base.py:17-18)."""

    def __init__(self, t):
        self.t, self.dtype = t, _np.dtype(t)

    def __call__(self, v):
        return self.t(v)

    def __hash__(self):
        return hash(self.dtype)

    def __eq__(self, o):
        return _np.dtype(getattr(o, "dtype", o)) == self.dtype


float16, float32, float64 = _ScalarType(_np.float16), _ScalarType(_np.float32), _ScalarType(_np.float64)
int8, int16, int32, int64 = (_ScalarType(t) for t in (_np.int8, _np.int16, _np.int32, _np.int64))
uint32, bool_ = _ScalarType(_np.uint32), _ScalarType(_np.bool_)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def set(self, v):
        out = _np.array(self.a, copy=True).view(ShimArray)
        out[self.idx] = v
        return out

    def add(self, v):
        out = _np.array(self.a, copy=True).view(ShimArray)
        out[self.idx] += v
        return out


class ShimArray(_np.ndarray):
    @property
    def at(self):
        return _At(self)


def _narrow(a):
    if a.dtype == _np.float64:
        return a.astype(_np.float32)
    if a.dtype == _np.int64:
        return a.astype(_np.int32)
    return a


def _wrap(a):
    return _np.asarray(a).view(ShimArray)


def array(obj, dtype=None):
    if isinstance(obj, (list, tuple)):
        obj = [_np.asarray(o) for o in obj]
    a = _np.array(obj, dtype=dtype)
    return _wrap(a if dtype is not None else _narrow(a))


def asarray(obj, dtype=None):
    if isinstance(obj, _np.ndarray) and dtype is None:
        return _wrap(_narrow(obj) if not isinstance(obj, ShimArray) else obj)
    return array(obj, dtype)


def zeros(shape, dtype=float32):
    return _wrap(_np.zeros(shape, dtype))


def ones(shape, dtype=float32):
    return _wrap(_np.ones(shape, dtype))


def zeros_like(a, dtype=None):
    return _wrap(_np.zeros_like(_np.asarray(a), dtype=dtype))


def ones_like(a, dtype=None):
    return _wrap(_np.ones_like(_np.asarray(a), dtype=dtype))


def eye(n, m=None, dtype=float32):
    return _wrap(_np.eye(n, m, dtype=dtype))


def arange(*a, dtype=None):
    return _wrap(_narrow(_np.arange(*a, dtype=dtype)))


def linspace(start, stop, num=50, dtype=None):
    return _wrap(_np.linspace(start, stop, num).astype(dtype or float32))


def _lift(name):
    f = getattr(_np, name)

    def g(*a, **k):
        r = f(*a, **k)
        return _wrap(_narrow(r)) if isinstance(r, _np.ndarray) else r
    g.__name__ = name
    return g


for _n in ("cos", "sin", "exp", "log", "sqrt", "tanh", "abs", "clip", "squeeze", "where", "sum", "mean", "any", "all",
           "isnan", "isinf", "isfinite", "argmax", "argmin", "argsort", "vstack", "hstack", "stack", "concatenate",
           "maximum", "minimum", "power", "square", "outer", "dot", "matmul", "diag", "cumsum", "expand_dims", "reshape",
           "transpose", "tile", "repeat", "round", "floor", "ceil", "sign", "prod", "max", "min", "mod", "remainder",
           "logical_and", "logical_or", "logical_not", "nan_to_num", "take", "flip", "sort", "unique", "nonzero",
           "full", "count_nonzero", "atleast_1d", "atleast_2d", "ravel", "log1p", "arctan", "arctan2", "linalg"):
    globals()[_n] = _lift(_n) if _n != "linalg" else _np.linalg
