"""``jax.lax`` stand-in: ``scan`` as a Python loop (test infrastructure only)."""
from . import tree_util as _tu
from . import numpy as _jnp


def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
    if xs is None:
        n = int(length)
    else:
        n = int(_tu.tree_leaves(xs)[0].shape[0])
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        x = None if xs is None else _tu.tree_map(lambda a: a[i], xs)
        carry, y = f(carry, x)
        ys[i] = y
    if n == 0:
        return carry, None
    stacked = _tu.tree_map(lambda *leaves: _jnp.stack(list(leaves)), ys[0], *ys[1:])
    return carry, stacked
