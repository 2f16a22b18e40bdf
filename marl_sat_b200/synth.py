"""Synthetic CNF generators (host side, NumPy) shared by bench.py and the tests.

``uniform_ksat`` draws SATLIB "uf"-style uniform random k-SAT: each clause has k distinct variables
drawn uniformly from n, each negated with probability 1/2, no satisfiability filtering
(SURVEY.md section 8d).  ``mixed_ksat`` draws clause widths uniformly in [kmin, kmax] and pads
with literal 0 up to kmax (BASELINE config 5).
"""
from __future__ import annotations

import numpy as np


def uniform_ksat(num_formulas: int, n: int, m: int, k: int = 3, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if k > n:
        raise ValueError("k distinct variables need k <= n")
    vars_ = rng.integers(0, n, size=(num_formulas, m, k), dtype=np.int16 if n < 2 ** 15 else np.int32)
    while True:   # re-draw rows with a repeated variable
        dup = np.zeros(vars_.shape[:2], dtype=bool)
        for i in range(k):
            for j in range(i + 1, k):
                dup |= vars_[:, :, i] == vars_[:, :, j]
        cnt = int(dup.sum())
        if cnt == 0:
            break
        vars_[dup] = rng.integers(0, n, size=(cnt, k), dtype=vars_.dtype)
    neg = rng.integers(0, 2, size=vars_.shape, dtype=np.int8).astype(bool)
    lits = vars_.astype(np.int32) + 1
    return np.where(neg, -lits, lits)


def mixed_ksat(num_formulas: int, n: int, m: int, kmin: int = 3, kmax: int = 7, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    full = uniform_ksat(num_formulas, n, m, kmax, seed=seed + 1)
    width = rng.integers(kmin, kmax + 1, size=(num_formulas, m, 1))
    keep = np.arange(kmax)[None, None, :] < width
    return np.where(keep, full, 0).astype(np.int32)


def uniform_ksat_torch(num_formulas: int, n: int, m: int, k: int = 3, seed: int = 0, device="cpu"):
    """Same distribution as ``uniform_ksat`` generated with torch on ``device`` (bench set-up of tens of
    thousands of formulas in milliseconds on the GPU).  Deterministic per (seed, device type); the values
    differ from the NumPy generator's."""
    import torch
    if k > n:
        raise ValueError("k distinct variables need k <= n")
    g = torch.Generator(device=device).manual_seed(seed)
    vars_ = torch.randint(0, n, (num_formulas, m, k), generator=g, device=device, dtype=torch.int32)
    while True:   # re-draw rows with a repeated variable
        dup = torch.zeros((num_formulas, m), dtype=torch.bool, device=device)
        for i in range(k):
            for j in range(i + 1, k):
                dup |= vars_[:, :, i] == vars_[:, :, j]
        cnt = int(dup.sum())
        if cnt == 0:
            break
        vars_[dup] = torch.randint(0, n, (cnt, k), generator=g, device=device, dtype=torch.int32)
    neg = torch.randint(0, 2, vars_.shape, generator=g, device=device, dtype=torch.int32).bool()
    lits = vars_ + 1
    return torch.where(neg, -lits, lits)


def mixed_ksat_torch(num_formulas: int, n: int, m: int, kmin: int = 3, kmax: int = 7, seed: int = 0, device="cpu"):
    import torch
    full = uniform_ksat_torch(num_formulas, n, m, kmax, seed=seed + 1, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    width = torch.randint(kmin, kmax + 1, (num_formulas, m, 1), generator=g, device=device)
    keep = torch.arange(kmax, device=device)[None, None, :] < width
    return torch.where(keep, full, torch.zeros_like(full))
