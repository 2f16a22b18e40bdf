"""``jax.random`` stand-in: legacy uint32[2] Threefry keys via oracle/threefry.py, which is pinned by
the Random123 and JAX-documentation known-answer vectors (test infrastructure only)."""
import numpy as _np

from oracle import threefry as _tf

from ._array import unwrap as _u, wrap as _w


def PRNGKey(seed):
    return _w(_tf.prng_key(int(seed)))


def split(key, num=2):
    return _w(_tf.split(_np.asarray(_u(key), dtype=_np.uint32), int(num)))


def randint(key, shape, minval, maxval, dtype=_np.int32):
    shape = tuple(shape)
    n = int(_np.prod(shape)) if shape else 1
    out = _tf.randint(_np.asarray(_u(key), dtype=_np.uint32), n, int(minval), int(maxval))
    return _w(out.reshape(shape).astype(dtype))
