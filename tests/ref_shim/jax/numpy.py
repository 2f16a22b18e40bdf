"""``jax.numpy`` stand-in on NumPy (test infrastructure only; see ../README.md)."""
from __future__ import annotations

import numpy as _np

from ._array import Array, unwrap as _u, wrap as _w, promote_inputs as _promote

ndarray = Array
int32, uint32, float32, bool_, int8, uint8 = _np.int32, _np.uint32, _np.float32, _np.bool_, _np.int8, _np.uint8
float64, int64 = _np.float32, _np.int32          # x64 disabled
newaxis = None
pi = _np.pi
inf = _np.inf


def _ul(seq):
    return [_u(s) for s in seq]


def array(x, dtype=None, copy=True):
    if isinstance(x, (list, tuple)):
        x = _np.array([_np.asarray(_u(e)) for e in x]) if len(x) and not _np.isscalar(x[0]) else _np.array(x)
    return _w(_np.array(_u(x), dtype=dtype))


def asarray(x, dtype=None):
    return array(x, dtype=dtype)


def zeros(shape, dtype=float32):
    return _w(_np.zeros(shape, dtype=dtype))


def ones(shape, dtype=float32):
    return _w(_np.ones(shape, dtype=dtype))


def full(shape, fill_value, dtype=None):
    return _w(_np.full(shape, _u(fill_value), dtype=dtype))


def zeros_like(x, dtype=None):
    return _w(_np.zeros_like(_u(x), dtype=dtype))


def ones_like(x, dtype=None):
    return _w(_np.ones_like(_u(x), dtype=dtype))


def arange(*a, dtype=None):
    return _w(_np.arange(*a, dtype=dtype))


def abs(x):
    return _w(_np.abs(_u(x)))


def where(c, x, y):
    x, y = _promote([x, y])
    return _w(_np.where(_u(c), x, y))


def any(x, axis=None, keepdims=False):
    return _w(_np.any(_u(x), axis=axis, keepdims=keepdims))


def all(x, axis=None, keepdims=False):
    return _w(_np.all(_u(x), axis=axis, keepdims=keepdims))


def sum(x, axis=None, keepdims=False, dtype=None):
    return _w(_np.sum(_u(x), axis=axis, keepdims=keepdims, dtype=dtype))


def mean(x, axis=None, keepdims=False):
    return _w(_np.mean(_u(x), axis=axis, keepdims=keepdims))


def var(x, axis=None):
    return _w(_np.var(_u(x), axis=axis))


def std(x, axis=None):
    return _w(_np.std(_u(x), axis=axis))


def maximum(a, b):
    a, b = _promote([a, b])
    return _w(_np.maximum(a, b))


def minimum(a, b):
    a, b = _promote([a, b])
    return _w(_np.minimum(a, b))


def logical_xor(a, b):
    return _w(_np.logical_xor(_u(a), _u(b)))


def logical_and(a, b):
    return _w(_np.logical_and(_u(a), _u(b)))


def logical_or(a, b):
    return _w(_np.logical_or(_u(a), _u(b)))


def logical_not(a):
    return _w(_np.logical_not(_u(a)))


def stack(xs, axis=0):
    return _w(_np.stack(_promote(list(xs)), axis=axis))


def concatenate(xs, axis=0):
    return _w(_np.concatenate(_promote(list(xs)), axis=axis))


def reshape(x, shape):
    return _w(_np.reshape(_u(x), shape))


def broadcast_to(x, shape):
    return _w(_np.array(_np.broadcast_to(_u(x), shape)))


def argmax(x, axis=None):
    return _w(_np.argmax(_u(x), axis=axis))


def unique(x, size=None, fill_value=None):
    """``jnp.unique`` with the static ``size``: sorted unique values, padded with ``fill_value``
    (default: the minimum value) or truncated to ``size``."""
    vals = _np.unique(_u(x))
    if size is not None:
        if vals.size >= size:
            vals = vals[:size]
        else:
            fv = vals.min() if fill_value is None else fill_value
            vals = _np.concatenate([vals, _np.full(size - vals.size, fv, dtype=vals.dtype)])
    return _w(vals)


def flatnonzero(x):
    return _w(_np.flatnonzero(_u(x)))


def copy(x):
    return _w(_np.array(_u(x), copy=True))


def sqrt(x):
    return _w(_np.sqrt(*_promote([x])))


def clip(x, a_min=None, a_max=None):
    return _w(_np.clip(_u(x), a_min, a_max))


def take_along_axis(x, idx, axis):
    return _w(_np.take_along_axis(_u(x), _u(idx), axis=axis))
