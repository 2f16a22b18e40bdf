"""NumPy restatement of the rollout auto-reset + RNG chain (test oracle / CPU baseline only).

Follows ``/root/reference/src/learners/mappo_gnn_sat_learner.py:383-480`` (one
``_env_step`` minus the policy) and ``/root/reference/src/runners/mappo_runner.py:289-295``
(initial reset).  Pinned by reference-generated fixtures (``tests/test_golden_env.py``; see ``oracle/__init__.py``).
"""
from __future__ import annotations

from dataclasses import fields, replace
from typing import Dict

import numpy as np

from . import threefry
from .sat_env import SATEnvOracle, SATState


def rollout_keys(rng: np.ndarray, num_envs: int, num_problems: int) -> Dict[str, np.ndarray]:
    """The per-step key chain of learner:397,416-417,426-434.

    rng -> (rng, act_key) -> (rng, step_key) -> (rng, prob_key, reset_key);
    ``new_problem_indices = randint(prob_key, (B,), 0, P)``;
    ``reset_keys = split(reset_key, B)``.
    """
    rng, act_key = threefry.split(rng)                       # learner:397
    rng, step_key = threefry.split(rng)                      # learner:416
    rng, prob_key, reset_key = threefry.split(rng, 3)        # learner:426
    idx = threefry.randint(prob_key, num_envs, 0, num_problems)   # learner:430
    reset_keys = threefry.split(reset_key, num_envs)         # learner:434
    return {"rng": rng, "act_key": act_key, "step_key": step_key, "prob_key": prob_key,
            "reset_key": reset_key, "new_problem_indices": idx, "reset_keys": reset_keys}


def initial_reset_inputs(key: np.ndarray, num_envs: int, num_problems: int):
    """runner:289-295: ``key, _rng = split(key)``; the *same* ``_rng`` feeds both
    ``randint`` and ``split``."""
    key, _rng = threefry.split(key)
    idx = threefry.randint(_rng, num_envs, 0, num_problems)
    reset_keys = threefry.split(_rng, num_envs)
    return key, idx, reset_keys


def _select(done: np.ndarray, new: np.ndarray, old: np.ndarray) -> np.ndarray:
    """``_reset_if_done`` (learner:445-451)."""
    mask = done.reshape(done.shape + (1,) * (old.ndim - 1))
    return np.where(mask, new, old)


def env_step_with_autoreset(env: SATEnvOracle, state: SATState, actions: np.ndarray,
                            problems_clauses: np.ndarray, new_problem_indices: np.ndarray,
                            reset_keys: np.ndarray):
    """One rollout step structured exactly like the reference: step every env,
    reset *every* env on its newly drawn problem, then select per leaf with
    ``done["__all__"]`` (learner:418-464).

    Returns ``(final_obs[B,A,D], final_state, reward[B,A], done_all[B], info)``
    where reward/done/info are the pre-reset values stored in the Transition
    (learner:467-478).
    """
    obs, nxt, rewards, dones, infos = env.step_env(None, state, actions)
    done_all = dones["__all__"]
    new_clauses = problems_clauses[new_problem_indices]                        # learner:431
    obs_r, state_r = env.reset(new_clauses, reset_keys)                        # learner:435
    leaves = {}
    for f in fields(SATState):
        old, new = getattr(nxt, f.name), getattr(state_r, f.name)
        if f.name == "action_mask":                                            # unbatched constant
            leaves[f.name] = old
        else:
            leaves[f.name] = _select(done_all, new, old)
    final_state = replace(nxt, **leaves)
    obs_a = np.stack([obs[a] for a in env.agents], axis=1)
    obs_ra = np.stack([obs_r[a] for a in env.agents], axis=1)
    final_obs = _select(done_all, obs_ra, obs_a)
    reward = np.stack([rewards[a] for a in env.agents], axis=-1)               # learner:471
    return final_obs, final_state, reward, done_all, infos


def rollout_T(env: SATEnvOracle, problems_clauses: np.ndarray, key0: np.ndarray, actions: np.ndarray,
              values: np.ndarray | None = None):
    """T rollout steps structured like the reference's ``lax.scan(_env_step, ...)`` (learner:383-495),
    starting from the runner's initial reset (runner:289-295), with the policy replaced by the given
    action table ``actions[T,B,A(,V)]`` (and ``values[T,B]``).  Returns ``(transition, final)``:

    ``transition`` -- the stacked ``Transition`` fields (learner:467-478): ``local_obs[T,B,A,D]`` and the
    GNN-input leaves are those of the state the policy acted on (**pre-step**); ``reward[T,B,A]``,
    ``global_done[T,B]`` and ``info`` are the **pre-reset** results of that step; ``agent_clause_masks`` /
    ``agent_neighbor_masks`` belong to the pre-step state.
    ``final`` -- the carry after the last step: state, obs, rng, plus the per-step ``act_key`` chain.
    """
    from .features import state_to_gnn_input
    T, B = actions.shape[0], actions.shape[1]
    P = problems_clauses.shape[0]
    rng, idx0, keys0 = initial_reset_inputs(np.asarray(key0, np.uint32), B, P)
    obs, state = env.reset(problems_clauses[idx0], keys0)
    obs = np.stack([obs[a] for a in env.agents], axis=1)
    tr = {k: [] for k in ("global_done", "action", "value", "reward", "local_obs", "gs_assignment",
                          "gs_clause_features", "info_solved", "info_num_unsatisfied", "info_episode_step",
                          "agent_clause_masks", "agent_neighbor_masks", "act_key")}
    initial = {"obs": obs, "state": state, "problem_idx": idx0, "reset_keys": keys0, "rng": rng}
    for t in range(T):
        gs = state_to_gnn_input(env, state)
        tr["local_obs"].append(obs)                                           # learner:473 (last_local_obs)
        tr["gs_assignment"].append(gs["assignment"])                          # learner:474 (last_global_state)
        tr["gs_clause_features"].append(gs["clause_features"])
        tr["agent_clause_masks"].append(state.agent_clause_masks)             # learner:386-387
        tr["agent_neighbor_masks"].append(state.agent_neighbor_masks)
        ks = rollout_keys(rng, B, P)
        rng = ks["rng"]
        tr["act_key"].append(ks["act_key"])
        obs, state, reward, done_all, infos = env_step_with_autoreset(
            env, state, actions[t], problems_clauses, ks["new_problem_indices"], ks["reset_keys"])
        tr["global_done"].append(done_all)                                    # learner:468
        tr["action"].append(actions[t])
        tr["value"].append(values[t] if values is not None else np.zeros((B,), np.float32))
        tr["reward"].append(reward)                                           # learner:471
        tr["info_solved"].append(infos["solved"])
        tr["info_num_unsatisfied"].append(infos["num_unsatisfied"])
        tr["info_episode_step"].append(infos["episode_step"])
    transition = {k: np.stack(v) for k, v in tr.items()}
    final = {"obs": obs, "state": state, "rng": rng, "initial": initial}
    return transition, final
