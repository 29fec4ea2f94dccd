"""A numpy stand-in for the small part of jax that the reference's OWN code needs (not diffrax, not jax.random's
generators): enough to import sdevries0/MultiTreeGP from /root/reference in an image without jax and to execute its
tree interpreter (genetic_programming.py: body_fun / foriloop / vmap_foriloop, node-library construction) and its
environment classes (drift, f_obs without noise, fitness_function, cond_fn_nan) on float32 numpy arrays.

Used only by tools/record_reference_shim_golden.py (test infrastructure).  Semantics kept: float32 default dtype,
weak python scalars (numpy >= 2 promotion), lax.switch index clamping, functional `.at[...]` updates, vmap as a loop.
Not kept: XLA's own float32 sin/cos/exp implementations (numpy's are used: a few ulp), jit, sharding, PRNG.
"""
import numpy as _np

from . import numpy as numpy  # noqa: F401
from . import lax as lax  # noqa: F401
from . import random as random  # noqa: F401
from .numpy import ShimArray as Array, _wrap

SHIM = True


def jit(f=None, **_kw):
    return f if f is not None else (lambda g: g)


def device_put(x, *_a, **_k):
    return x


def devices(_kind=None):
    return [object()]


def value_and_grad(_f, **_kw):
    def no(*_a, **_k):
        raise NotImplementedError("refshim: no autodiff")
    return no


grad = value_and_grad


def _axis_list(in_axes, n):
    if isinstance(in_axes, (list, tuple)):
        return list(in_axes)
    return [in_axes] * n


def _take(a, ax, i):
    if ax is None:
        return a
    if isinstance(a, (tuple, list)):
        return type(a)(_take(x, ax, i) for x in a)
    return _wrap(_np.take(_np.asarray(a), i, axis=ax))


def _size(a, ax):
    if isinstance(a, (tuple, list)):
        return _size(a[0], ax)
    return _np.asarray(a).shape[ax]


def _stack(outs):
    if isinstance(outs[0], (tuple, list)):
        return type(outs[0])(_stack([o[k] for o in outs]) for k in range(len(outs[0])))
    return _wrap(_np.stack([_np.asarray(o) for o in outs]))


def vmap(f, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = _axis_list(in_axes, len(args))
        n = next(_size(a, ax) for a, ax in zip(args, axes) if ax is not None)
        return _stack([f(*[_take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)])
    return mapped
