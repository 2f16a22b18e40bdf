"""NumPy-backed stand-in for the parts of ``jax`` the reference uses (TEST INFRASTRUCTURE ONLY --
see ../README.md).  Everything executes eagerly; ``jit`` is the identity and ``vmap`` is a loop."""
from __future__ import annotations

from . import numpy, nn, random, lax, tree_util          # noqa: F401
from . import numpy as _jnp

Array = numpy.ndarray
__version__ = "0.4.29-numpy-shim"


def jit(fn=None, static_argnums=None, static_argnames=None, **_):
    if fn is None:
        return lambda f: f
    return fn


def device_get(x):
    return x


def vmap(fn, in_axes=0, out_axes=0):
    """Loop over the mapped axis and stack every output leaf (axis 0 only)."""
    def wrapped(*args, **kwargs):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args), "in_axes does not match the positional arguments"
        assert all(a in (0, None) for a in axes) and out_axes == 0
        size = None
        for a, ax in list(zip(args, axes)) + [(v, 0) for v in kwargs.values()]:
            if ax is None:
                continue
            leaves = tree_util.tree_leaves(a)
            if leaves:
                size = int(numpy.asarray(leaves[0]).shape[0])
                break
        assert size is not None, "vmap needs at least one mapped array argument"
        outs = []
        for i in range(size):
            call_args = [a if ax is None else tree_util.tree_map(lambda x: numpy.asarray(x)[i], a)
                         for a, ax in zip(args, axes)]
            call_kwargs = {k: tree_util.tree_map(lambda x: numpy.asarray(x)[i], v) for k, v in kwargs.items()}
            outs.append(fn(*call_args, **call_kwargs))
        return tree_util.tree_map(lambda *leaves: _jnp.stack([_jnp.asarray(l) for l in leaves]), outs[0], *outs[1:])
    return wrapped


def tree_map(fn, tree, *rest):
    return tree_util.tree_map(fn, tree, *rest)


class _Debug:
    @staticmethod
    def print(*a, **k):
        pass


debug = _Debug()
