"""``jax.tree_util`` stand-in: dicts (sorted keys), lists, tuples, dataclasses and None are nodes
(test infrastructure only)."""
import dataclasses as _dc


def _is_node(t):
    return isinstance(t, (dict, list, tuple)) or t is None or (_dc.is_dataclass(t) and not isinstance(t, type))


def tree_flatten(t):
    leaves = []

    def rec(x):
        if x is None:
            return ("none",)
        if isinstance(x, dict):
            keys = sorted(x.keys())
            return ("dict", keys, [rec(x[k]) for k in keys])
        if isinstance(x, (list, tuple)):
            return ("seq", type(x), [rec(e) for e in x])
        if _dc.is_dataclass(x) and not isinstance(x, type):
            names = [f.name for f in _dc.fields(x)]
            return ("dc", type(x), names, [rec(getattr(x, n)) for n in names])
        leaves.append(x)
        return ("leaf",)
    return leaves, rec(t)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        kind = d[0]
        if kind == "none":
            return None
        if kind == "leaf":
            return next(it)
        if kind == "dict":
            return {k: rec(c) for k, c in zip(d[1], d[2])}
        if kind == "seq":
            vals = [rec(c) for c in d[2]]
            return d[1](vals) if d[1] in (list, tuple) else d[1](*vals)
        return d[1](**{n: rec(c) for n, c in zip(d[2], d[3])})
    return rec(treedef)


def tree_leaves(t):
    return tree_flatten(t)[0]


def tree_map(fn, tree, *rest):
    leaves, treedef = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(treedef, [fn(*xs) for xs in zip(leaves, *others)])
