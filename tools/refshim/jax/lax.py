"""jax.lax control flow as python control flow (fori_loop, switch with index clamping, select, cond, scan)."""
import numpy as _np
from .numpy import _wrap


def fori_loop(lo, hi, body, init):
    v = init
    for i in range(int(lo), int(hi)):
        v = body(i, v)
    return v


def switch(index, branches, *operands):
    i = int(index)
    i = 0 if i < 0 else (len(branches) - 1 if i >= len(branches) else i)  # lax.switch clamps the index
    return branches[i](*operands)


def select(pred, on_true, on_false):
    return _wrap(_np.where(pred, on_true, on_false)) if _np.ndim(pred) else (on_true if bool(pred) else on_false)


def cond(pred, t, f, *operands):
    return t(*operands) if bool(pred) else f(*operands)


def bitcast_convert_type(val, new_type):
    return _np.asarray(val).view(_np.dtype(getattr(new_type, 'dtype', new_type)))


def stop_gradient(x):
    return x


def scan(f, init, xs, length=None):
    """lax.scan over the leading axis of xs (an array or a tuple of arrays): (carry, stacked ys)."""
    from . import _take, _size, _stack  # the vmap helpers
    n = _size(xs, 0) if xs is not None else int(length)
    carry, ys = init, []
    for i in range(n):
        carry, y = f(carry, _take(xs, 0, i) if xs is not None else None)
        ys.append(y)
    return carry, _stack(ys)
