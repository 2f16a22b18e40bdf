"""ndarray subclass with the JAX behaviours the reference relies on (test infrastructure only).

* type promotion with x64 disabled: results are never float64 / int64; an integer operand meeting a
  floating operand (array or Python scalar) is converted to float32 *before* the operation, so the
  arithmetic itself happens in float32 like XLA's;
* gather semantics of ``x[idx]`` for integer indices: negative indices wrap once, indices that are
  still out of range are clamped (``jax.numpy`` default ``mode="fill"/"clip"`` for gathers = clip);
* functional updates ``x.at[idx].set(v)`` / ``.add(v)`` (scatter: negative indices wrap, out-of-range
  updates are dropped).
"""
from __future__ import annotations

import numbers

import numpy as np


def _narrow_dtype(dt: np.dtype) -> np.dtype:
    if dt == np.float64:
        return np.dtype(np.float32)
    if dt == np.int64:
        return np.dtype(np.int32)
    if dt == np.uint64:
        return np.dtype(np.uint32)
    return dt


def wrap(x):
    """NumPy result -> Array with 32-bit dtypes (also turns NumPy scalars into 0-d arrays)."""
    if isinstance(x, tuple):
        return tuple(wrap(v) for v in x)
    if isinstance(x, (np.ndarray, np.generic)):
        a = np.asarray(x)
        nd = _narrow_dtype(a.dtype)
        if nd != a.dtype:
            a = a.astype(nd)
        return a.view(Array)
    return x


def unwrap(x):
    if isinstance(x, Array):
        return np.asarray(x)
    return x


def _is_float(x) -> bool:
    if isinstance(x, np.ndarray) or isinstance(x, np.generic):
        return np.issubdtype(x.dtype, np.floating)
    return isinstance(x, float)


def _is_intlike(x) -> bool:
    if isinstance(x, np.ndarray) or isinstance(x, np.generic):
        return np.issubdtype(x.dtype, np.integer) or x.dtype == np.bool_
    return False


def promote_inputs(inputs):
    """JAX lattice restricted to what the reference uses: {bool, int32, uint32} + float -> float32."""
    ins = [unwrap(i) for i in inputs]
    if any(_is_float(i) for i in ins):
        out = []
        for i in ins:
            if _is_intlike(i):
                out.append(np.asarray(i).astype(np.float32))
            elif isinstance(i, (np.ndarray, np.generic)) and i.dtype == np.float64:
                out.append(np.asarray(i).astype(np.float32))
            elif isinstance(i, float):
                out.append(np.float32(i))
            elif isinstance(i, numbers.Integral) and not isinstance(i, bool):
                out.append(np.float32(i))
            else:
                out.append(i)
        return out
    return ins


def _norm_int_index(ix, size):
    a = np.asarray(ix)
    if a.dtype == np.bool_ or not np.issubdtype(a.dtype, np.integer):
        return ix
    a = a.astype(np.int64)
    a = np.where(a < 0, a + size, a)          # wrap once
    return np.clip(a, 0, max(size - 1, 0))    # then clamp (JAX gather)


def _gather_index(idx, shape):
    """Normalise integer (array) indices of a NumPy-style index expression the way JAX gathers do."""
    tup = idx if isinstance(idx, tuple) else (idx,)
    tup = tuple(unwrap(t) for t in tup)
    n_real = sum(1 for t in tup if t is not None and t is not Ellipsis)
    out, axis = [], 0
    for t in tup:
        if t is None:
            out.append(t)
        elif t is Ellipsis:
            out.append(t)
            axis += len(shape) - n_real
        elif isinstance(t, slice):
            out.append(t)
            axis += 1
        elif isinstance(t, np.ndarray) and t.dtype == np.bool_:
            out.append(t)
            axis += t.ndim
        elif isinstance(t, (numbers.Integral, np.ndarray, np.generic, list)):
            out.append(_norm_int_index(t, shape[axis]))
            axis += 1
        else:
            out.append(t)
            axis += 1
    return tuple(out) if isinstance(idx, tuple) else out[0]


class _AtIndexer:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _prep(self):
        idx = self.idx
        tup = idx if isinstance(idx, tuple) else (idx,)
        tup = tuple(unwrap(t) for t in tup)
        return tup if isinstance(idx, tuple) else tup[0]

    def set(self, value):
        out = np.array(np.asarray(self.arr), copy=True)
        out[self._prep()] = np.asarray(unwrap(value)).astype(out.dtype) if not np.isscalar(value) else value
        return wrap(out)

    def add(self, value):
        out = np.array(np.asarray(self.arr), copy=True)
        v = np.asarray(unwrap(value)).astype(out.dtype)
        np.add.at(out, self._prep(), v)
        return wrap(out)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndexer(self.arr, idx)


class Array(np.ndarray):
    __array_priority__ = 1000

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = promote_inputs(inputs)
        if out is not None:
            kwargs["out"] = tuple(unwrap(o) for o in out)
        res = getattr(ufunc, method)(*ins, **kwargs)
        return wrap(res)

    def __array_function__(self, func, types, args, kwargs):
        # route every NumPy function through plain ndarrays and re-wrap
        def conv(v):
            if isinstance(v, Array):
                return np.asarray(v)
            if isinstance(v, (list, tuple)):
                return type(v)(conv(e) for e in v)
            if isinstance(v, dict):
                return {k: conv(e) for k, e in v.items()}
            return v
        res = func(*conv(args), **conv(kwargs))
        if isinstance(res, list):
            return [wrap(r) for r in res]
        return wrap(res)

    def __getitem__(self, idx):
        res = np.asarray(self)[_gather_index(idx, self.shape)]
        return wrap(res)

    def __iter__(self):
        for i in range(self.shape[0]):
            yield self[i]

    @property
    def at(self):
        return _At(self)

    def astype(self, dtype, *a, **k):
        return wrap(np.asarray(self).astype(dtype, *a, **k))

    def __hash__(self):
        return id(self)

    def __bool__(self):
        return bool(np.asarray(self))

    def __int__(self):
        return int(np.asarray(self))

    def __float__(self):
        return float(np.asarray(self))

    def __index__(self):
        return int(np.asarray(self))
