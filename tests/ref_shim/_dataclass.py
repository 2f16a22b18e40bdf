"""Shared dataclass decorator for the chex / flax.struct stand-ins (test infrastructure only):
keyword-only frozen dataclass with ``.replace``; instances are pytree nodes (jax/tree_util.py)."""
import dataclasses


def dataclass(cls=None, frozen=True, **_):
    def deco(c):
        c = dataclasses.dataclass(c, frozen=frozen, kw_only=True, eq=False)
        c.replace = lambda self, **kw: dataclasses.replace(self, **kw)
        return c
    return deco if cls is None else deco(cls)
