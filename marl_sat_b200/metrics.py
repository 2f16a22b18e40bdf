"""Rollout metrics (next-tier row): the reductions of
``/root/reference/src/learners/mappo_gnn_sat_learner.py:661-686`` over a ``[T, B]`` rollout in one pass
(``msat_rollout_metrics``).  With ``torch.distributed`` initialised the five float64 sums are
all-reduced so every rank reports the global metrics."""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib
from .env import _ptr, _stream_ptr


def rollout_metric_sums(reward: torch.Tensor, done: torch.Tensor, solved: torch.Tensor,
                        num_unsatisfied: torch.Tensor, episode_step: torch.Tensor) -> torch.Tensor:
    """float64[5] = {sum team reward, #finished, #solved at finish, sum unsat at finish, sum steps of solved}."""
    lib = _lib.load()
    T, B = done.shape[0], done.shape[1]
    u8 = lambda t: (t.view(torch.uint8) if t.dtype == torch.bool else t).contiguous()
    sums = torch.zeros(5, dtype=torch.float64, device=reward.device)
    _lib.check(lib.msat_rollout_metrics(_ptr(reward), reward.stride(0), reward.stride(1), _ptr(u8(done)),
                                        _ptr(u8(solved)), _ptr(num_unsatisfied.contiguous()),
                                        _ptr(episode_step.contiguous()), T, B, _ptr(sums),
                                        _stream_ptr(reward.device)), "msat_rollout_metrics")
    return sums


def metrics_from_sums(sums, num_envs_global: int) -> Dict[str, float]:
    """learner:664-686 from the five (global) sums."""
    total_reward, finished, n_solved, unsat, steps = (float(x) for x in sums.tolist())
    return {"mean_episodic_return": total_reward / num_envs_global,         # learner:666-668
            "solve_rate": n_solved / max(finished, 1.0),                   # learner:674
            "avg_unsatisfied_clauses": unsat / max(finished, 1.0),         # learner:680
            "avg_steps_to_solve": steps / max(n_solved, 1.0)}              # learner:686


def rollout_metrics(reward, done, solved, num_unsatisfied, episode_step, num_envs_global=None, group=None) -> Dict[str, float]:
    """``mean_episodic_return``, ``solve_rate``, ``avg_unsatisfied_clauses``, ``avg_steps_to_solve``
    (learner:664-686).  ``reward`` is ``[T,B,A]`` (agent 0 read) or ``[T,B]``.  With ``torch.distributed``
    initialised the five sums are all-reduced, so every rank reports the metrics of the global batch."""
    from .gae import _world_size, allreduce_stats
    sums = allreduce_stats(rollout_metric_sums(reward, done, solved, num_unsatisfied, episode_step), group)
    B = done.shape[1]
    world = _world_size(group)
    if world > 1:
        B = num_envs_global if num_envs_global is not None else B * world
    return metrics_from_sums(sums, B)
