"""Minimal mirrors of the jaxmarl 0.0.7 ``spaces`` classes the reference SATEnv builds
(src/envs/multi_agent_sat_env.py:68-84): ``Discrete``, ``MultiDiscrete`` and ``Box``.
Only the attributes callers read (``n``, ``num_categories``, ``shape``, ``dtype``, ``low``,
``high``) and ``contains`` / ``sample`` are provided; sampling uses torch, not JAX keys.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch


class Space:
    shape: Tuple[int, ...] = ()
    dtype = torch.int32


class Discrete(Space):
    def __init__(self, num_categories: int, dtype=torch.int32):
        assert num_categories >= 0
        self.n = int(num_categories)
        self.shape = ()
        self.dtype = dtype

    def sample(self, generator: torch.Generator | None = None, batch: Sequence[int] = (), device=None):
        return torch.randint(0, self.n, tuple(batch), generator=generator, device=device, dtype=torch.int32)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(np.all((x >= 0) & (x < self.n)))

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    def __init__(self, num_categories: Sequence[int]):
        self.num_categories = np.asarray(num_categories, dtype=np.int64)
        self.shape = (len(num_categories),)
        self.dtype = torch.int32

    def sample(self, generator: torch.Generator | None = None, batch: Sequence[int] = (), device=None):
        hi = torch.as_tensor(self.num_categories, device=device)
        u = torch.rand(tuple(batch) + self.shape, generator=generator, device=device)
        return (u * hi).floor().to(torch.int32)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(np.all((x >= 0) & (x < self.num_categories)))

    def __repr__(self):
        return f"MultiDiscrete({self.num_categories.tolist()})"


class Box(Space):
    """Declared float32 in the reference (env:84) although observations are int32 (env:390-396)."""

    def __init__(self, low: float, high: float, shape: Tuple[int, ...], dtype=torch.float32):
        self.low, self.high = low, high
        self.shape = tuple(shape)
        self.dtype = dtype

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(np.all((x >= self.low) & (x <= self.high)))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape})"
