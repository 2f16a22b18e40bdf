"""NumPy restatement of the JAX 0.4.29 Threefry-2x32 PRNG (test oracle only).

The reference pins ``jax==0.4.29`` (requirements.txt:100,109) and uses legacy
``uint32[2]`` keys with the default ``jax_threefry_partitionable=False``.
JAX is a third-party dependency that is absent from /root/reference and not
installable offline, so this file restates the published algorithm
(``jax/_src/prng.py``: ``threefry2x32``, ``threefry_split``,
``threefry_random_bits``; ``jax/_src/random.py``: ``_randint``) and is anchored
on the reference's call sites:

* ``jax.random.randint(key, (n,), 0, 2)``            env:162
* ``jax.random.split(rng)`` / ``split(rng, 3)``       learner:397,416,426
* ``jax.random.split(step_key, NUM_ENVS)``            learner:417,434
* ``jax.random.randint(prob_key, (B,), 0, P)``        learner:430; runner:291

Known-answer vectors (Random123 + JAX docs) are checked in
tests/test_threefry_kat.py.
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)
_PARITY = _U32(0x1BD11BDA)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return ((x << _U32(r)) | (x >> _U32(32 - r))).astype(_U32)


def threefry2x32(k0, k1, x0, x1):
    """20-round Threefry-2x32 block function, vectorised over the counters.

    ``k0, k1`` scalars (or arrays broadcastable with the counters); ``x0, x1``
    uint32 arrays.  Returns the two output word arrays.
    """
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=_U32)
        k1 = np.asarray(k1, dtype=_U32)
        x0 = np.array(x0, dtype=_U32, copy=True)
        x1 = np.array(x1, dtype=_U32, copy=True)
        ks = (k0, k1, (k0 ^ k1 ^ _PARITY).astype(_U32))
        x0 = (x0 + ks[0]).astype(_U32)
        x1 = (x1 + ks[1]).astype(_U32)
        for g in range(5):
            for r in (_ROT_A if g % 2 == 0 else _ROT_B):
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(_U32)
            x0 = (x0 + ks[(g + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(g + 2) % 3] + _U32(g + 1)).astype(_U32)
        return x0, x1


def threefry_2x32(key, counts) -> np.ndarray:
    """``jax._src.prng.threefry_2x32``: hash a flat counter array with ``key``.

    Odd-length inputs are padded with one zero; the first half of the counters
    feeds word 0 and the second half word 1; the outputs are concatenated and
    truncated back to the input length.
    """
    key = np.asarray(key, dtype=_U32)
    counts = np.asarray(counts, dtype=_U32).ravel()
    n = counts.size
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, _U32)])
    half = counts.size // 2
    o0, o1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
    return np.concatenate([o0, o1])[:n]


def prng_key(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey(seed)`` for 0 <= seed < 2**32 -> ``[0, seed]``."""
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def split(key, num: int = 2) -> np.ndarray:
    """``jax.random.split(key, num)`` -> uint32[num, 2]."""
    return threefry_2x32(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits32(key, n: int) -> np.ndarray:
    """``_random_bits(key, 32, (n,))`` -> uint32[n]."""
    return threefry_2x32(key, np.arange(n, dtype=_U32))


def randint(key, n: int, minval: int, maxval: int) -> np.ndarray:
    """``jax.random.randint(key, (n,), minval, maxval)`` for int32.

    Two independent 32-bit draws (from the two halves of ``split(key)``) are
    combined as ``((hi % span) * mult + (lo % span)) % span`` in wrapping
    uint32 arithmetic, ``mult = (2**16 % span)**2 % span``.
    """
    k1, k2 = split(key)
    hi = random_bits32(k1, n)
    lo = random_bits32(k2, n)
    span = _U32(maxval - minval) if maxval > minval else _U32(1)
    with np.errstate(over="ignore"):
        mult = _U32(_U32(1 << 16) % span)
        mult = _U32((mult * mult) % span)
        off = ((hi % span) * mult + (lo % span)).astype(_U32) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def uniform01(key) -> np.float32:
    """``jax.random.uniform(key)`` (scalar) -- used only as a published KAT."""
    bits = random_bits32(key, 1)[0]
    f = np.array([(bits >> _U32(9)) | _U32(0x3F800000)], dtype=_U32).view(np.float32)[0]
    return np.float32(f - np.float32(1.0))


# --- batched helpers used by the env/rollout oracle -------------------------

def split_many(keys: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """``vmap(lambda k: split(k))`` over uint32[B,2] keys -> two uint32[B,2]."""
    keys = np.asarray(keys, dtype=_U32)
    k0, k1 = keys[:, 0], keys[:, 1]
    # counters [0,1,2,3] -> blocks (0,2) and (1,3); outputs [a0,a1,b0,b1]
    a0, b0 = threefry2x32(k0, k1, np.zeros_like(k0), np.full_like(k0, 2))
    a1, b1 = threefry2x32(k0, k1, np.ones_like(k0), np.full_like(k0, 3))
    return np.stack([a0, a1], axis=1), np.stack([b0, b1], axis=1)


def randint01_many(keys: np.ndarray, n: int) -> np.ndarray:
    """``vmap(lambda k: randint(k, (n,), 0, 2))`` -> int32[B, n].

    For span 2 the multiplier is 0, so the result is bit 0 of the draw made
    with the *second* sub-key (env:162).
    """
    _, k2 = split_many(keys)
    npad = n + (n & 1)
    half = npad // 2
    lo_ctr = np.arange(half, dtype=_U32)[None, :]
    hi_ctr = (np.arange(half, dtype=_U32) + _U32(half))[None, :]
    if n & 1:
        hi_ctr = hi_ctr.copy()
        hi_ctr[0, -1] = 0  # the zero pad
    o0, o1 = threefry2x32(k2[:, 0:1], k2[:, 1:2],
                          np.broadcast_to(lo_ctr, (keys.shape[0], half)),
                          np.broadcast_to(hi_ctr, (keys.shape[0], half)))
    bits = np.concatenate([o0, o1], axis=1)[:, :n]
    return (bits & _U32(1)).astype(np.int32)
