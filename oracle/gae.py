"""NumPy float32 restatement of the MAPPO GAE scan + advantage normalisation
(test oracle / CPU baseline only).

Follows ``/root/reference/src/learners/mappo_gnn_sat_learner.py:504-532``.
Pinned: bit-identical to the reference's own ``_calculate_gae`` run under the NumPy stand-ins
(``tests/test_golden_env.py``); tolerance on the CUDA path is 1e-5 relative as stated by
BASELINE.json's north_star.
"""
from __future__ import annotations

import numpy as np


def calculate_gae(reward: np.ndarray, done: np.ndarray, value: np.ndarray, last_val: np.ndarray,
                  gamma: float, gae_lambda: float):
    """reward f32[T,B,A] (agent 0 is read, learner:514) or f32[T,B]; done bool[T,B];
    value f32[T,B]; last_val f32[B] -> (advantages f32[T,B], targets f32[T,B])."""
    f32 = np.float32
    team = reward[..., 0] if reward.ndim == 3 else reward
    team = team.astype(f32)
    value = value.astype(f32)
    T, B = value.shape
    g = f32(gamma)
    gl = f32(gamma * gae_lambda)          # Python-double product rounded once
    gae = np.zeros((B,), f32)
    next_value = last_val.astype(f32)
    adv = np.empty((T, B), f32)
    for t in range(T - 1, -1, -1):        # reverse scan, learner:519-525
        nt = (1 - done[t].astype(np.int32)).astype(f32)
        delta = team[t] + g * next_value * nt - value[t]       # learner:515
        gae = delta + gl * nt * gae                              # learner:516
        adv[t] = gae
        next_value = value[t]
    return adv, adv + value                                      # learner:526


def normalize_advantages(adv: np.ndarray) -> np.ndarray:
    """learner:530-532: global mean / population std over all T*B, +1e-8 outside."""
    mean = adv.mean(dtype=np.float64)
    std = adv.std(dtype=np.float64)
    return ((adv - np.float32(mean)) / (np.float32(std) + np.float32(1e-8))).astype(np.float32)
