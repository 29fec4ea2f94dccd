"""jax.random is third-party to this check: keys are opaque, every sampler returns zeros (noise-free f_obs)."""
import numpy as _np
from .numpy import _wrap, float32


def PRNGKey(seed):
    return _wrap(_np.array([0, seed], dtype=_np.uint32))


key = PRNGKey


def split(k, num=2):
    return _wrap(_np.stack([_np.asarray(k)] * num))


def fold_in(k, _data):
    return k


def normal(_k, shape=(), dtype=float32):
    return _wrap(_np.zeros(shape, dtype))


def uniform(_k, shape=(), dtype=float32, minval=0.0, maxval=1.0):
    return _wrap(_np.zeros(shape, dtype) + _np.asarray(minval, dtype))
