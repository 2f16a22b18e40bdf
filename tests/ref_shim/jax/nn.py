"""``jax.nn`` stand-in (test infrastructure only)."""
import numpy as _np

from ._array import unwrap as _u, wrap as _w


def one_hot(x, num_classes, axis=-1, dtype=_np.float32):
    """Out-of-range classes (e.g. -1) give an all-zero row, like jax.nn.one_hot."""
    assert axis == -1
    x = _np.asarray(_u(x))
    return _w((x[..., None] == _np.arange(num_classes)).astype(dtype))


def relu(x):
    return _w(_np.maximum(_u(x), 0))
