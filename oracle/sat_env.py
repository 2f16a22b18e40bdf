"""NumPy restatement of the reference ``SATEnv`` (test oracle / CPU baseline only).

Follows ``/root/reference/src/envs/multi_agent_sat_env.py`` function by
function; every array carries a leading batch axis ``B`` where the reference
relies on ``jax.vmap`` (runner:137, learner:418).  Pinned: it reproduces, bit for
bit, the fixtures ``tests/golden/env_*.npz`` that the reference's own unmodified
source produced (``tests/golden/make_golden_env.py``; see ``oracle/__init__.py``).

JAX indexing semantics that matter and are reproduced here:
* negative gather indices wrap once (``x[-1]`` is the last element), indices
  that are still out of range are clamped (env:139, env:239);
* ``jax.nn.one_hot`` of an out-of-range class (``-1``) is all zeros (env:243).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import threefry


@dataclass(frozen=True)
class SATState:
    """Batched mirror of ``SATState`` (env:13-24) + jaxmarl ``State`` (done, step)."""
    variable_assignments: np.ndarray      # i32[B, n]
    clauses_satisfied_status: np.ndarray  # bool[B, m]
    num_unsatisfied: np.ndarray           # i32[B]
    step: np.ndarray                      # i32[B]
    done: np.ndarray                      # bool[B, A]
    clauses: np.ndarray                   # i32[B, m, k]
    agent_clause_masks: np.ndarray        # i32[B, A, m]  in {1,-1}
    agent_neighbor_masks: np.ndarray      # i32[B, A, n]  in {1,-1}
    literal_to_agent_idx: np.ndarray      # i32[B, m, k]
    action_mask: np.ndarray               # bool[A, V]


def _jax_gather_index(idx: np.ndarray, size: int) -> np.ndarray:
    """Index normalisation of ``x[idx]`` in JAX: wrap negatives once, then clamp."""
    idx = np.where(idx < 0, idx + size, idx)
    return np.clip(idx, 0, size - 1)


def create_agent_groups(num_vars: int, vars_per_agent: Optional[int]) -> Dict[str, List[int]]:
    """env:294-338.  Manual: ceil(n/vpa) agents; auto: n/4 agents if 4 | n else
    max(2, int(sqrt(n))) agents; contiguous ranges, first ``n % A`` one larger."""
    if vars_per_agent is not None:
        num_agents = math.ceil(num_vars / vars_per_agent)
    else:
        factors = set()
        for i in range(1, int(math.sqrt(num_vars)) + 1):   # _find_factors, env:286-293
            if num_vars % i == 0:
                factors.add(i)
                factors.add(num_vars // i)
        candidates = [f for f in sorted(factors) if 4 <= f <= 4]
        if candidates:
            num_agents = num_vars // max(candidates)
        else:
            num_agents = max(2, int(math.sqrt(num_vars)))
    base, rem = divmod(num_vars, num_agents)
    groups, cur = {}, 0
    for i in range(num_agents):
        size = base + 1 if i < rem else base
        groups[f"agent_{i}"] = list(range(cur, cur + size))
        cur += size
    return groups


class SATEnvOracle:
    """Batched NumPy mirror of ``SATEnv`` (env:28-411)."""

    def __init__(self, num_vars, num_clauses, max_steps: int, vars_per_agent: Optional[int] = None,
                 action_mode: int = 0, r_clause: float = 0.02, r_sat: float = 1.0, gamma: float = 0.99,
                 reward_mode: str = "sparse"):
        self.reward_mode = reward_mode                                  # "shaped": the commented variant env:201-223
        self.num_vars = num_vars
        self.num_clauses = num_clauses
        self.agent_groups = create_agent_groups(num_vars, vars_per_agent)
        self.agents = list(self.agent_groups.keys())
        self.num_agents = len(self.agents)
        self.agent_to_idx = {a: i for i, a in enumerate(self.agents)}
        self.r_clause, self.r_sat, self.gamma = r_clause, r_sat, gamma   # stored, unused (env:40-42)
        self.action_mode = action_mode
        self.max_vars_per_agent = max(len(v) for v in self.agent_groups.values())
        A, V = self.num_agents, self.max_vars_per_agent
        self.agent_vars = np.full((A, V), -1, dtype=np.int32)           # env:61
        self.action_mask = np.zeros((A, V), dtype=bool)                 # env:62
        for i, a in enumerate(self.agents):
            vs = self.agent_groups[a]
            self.agent_vars[i, :len(vs)] = vs
            self.action_mask[i, :len(vs)] = True
        self.obs_dim = 2 * num_vars + num_clauses                       # env:340-343
        self.max_steps = max_steps
        self.variable_to_agent_idx = np.full((num_vars,), -1, dtype=np.int32)   # env:92-97
        for i, a in enumerate(self.agents):
            self.variable_to_agent_idx[self.agent_groups[a]] = i

    # -- env:99-128 ---------------------------------------------------------
    def compute_observation_maps(self, clauses: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        B, m, k = clauses.shape
        n, A = self.num_vars, self.num_agents
        vi = np.abs(clauses) - 1                                        # env:100 (0 -> -1)
        acm = np.empty((B, A, m), dtype=np.int32)
        anm = np.empty((B, A, n), dtype=np.int32)
        rows = np.arange(B)[:, None]
        for a in range(A):
            row = self.agent_vars[a]                                    # padded with -1
            # env:106-107: equality against the *padded* row (so -1 == -1 matches)
            related = np.isin(vi, row).any(axis=2)                      # [B, m]
            acm[:, a] = np.where(related, 1, -1)                        # env:112
            # env:116-121: variables that occur in a related clause
            rel_vars = np.where(related[:, :, None], vi, -1).reshape(B, m * k)
            present = np.zeros((B, n + 1), dtype=bool)                  # slot n collects the -1s
            present[rows, np.where(rel_vars < 0, n, rel_vars)] = True
            is_related = present[:, :n]
            is_own = np.isin(np.arange(n), row)                         # env:120
            anm[:, a] = np.where(is_related & ~is_own[None, :], 1, -1)  # env:123-124
        return acm, anm

    # -- env:130-156 --------------------------------------------------------
    def calculate_satisfaction(self, assign: np.ndarray, clauses: np.ndarray):
        n = self.num_vars
        vi = _jax_gather_index(np.abs(clauses) - 1, n)                  # env:135,139
        B = assign.shape[0]
        a_lit = np.take_along_axis(assign, vi.reshape(B, -1), axis=1).reshape(clauses.shape)
        truth = ((clauses > 0) & (a_lit == 1)) | ((clauses < 0) & (a_lit == 0))   # env:141-144
        status = truth.any(axis=2)                                      # env:151
        num_unsat = (~status).sum(axis=1).astype(np.int32)              # env:154
        return status, num_unsat

    # -- env:158-181 --------------------------------------------------------
    def reset(self, problem_clauses: np.ndarray, keys: np.ndarray):
        clauses = np.asarray(problem_clauses, dtype=np.int32)
        keys = np.asarray(keys, dtype=np.uint32)
        squeeze = clauses.ndim == 2
        if squeeze:
            clauses, keys = clauses[None], keys[None]
        B = clauses.shape[0]
        l2a = self.variable_to_agent_idx[_jax_gather_index(np.abs(clauses) - 1, self.num_vars)]  # env:160
        acm, anm = self.compute_observation_maps(clauses)               # env:161
        assign = threefry.randint01_many(keys, self.num_vars)           # env:162
        status, num_unsat = self.calculate_satisfaction(assign, clauses)
        state = SATState(
            variable_assignments=assign, clauses_satisfied_status=status, num_unsatisfied=num_unsat,
            step=np.zeros((B,), np.int32), done=np.zeros((B, self.num_agents), bool), clauses=clauses,
            agent_clause_masks=acm, agent_neighbor_masks=anm, literal_to_agent_idx=l2a.astype(np.int32),
            action_mask=self.action_mask)
        return self.get_obs(state), state

    # -- env:225-284 --------------------------------------------------------
    def step_env(self, key, state: SATState, actions_array: np.ndarray):
        del key                                                         # unused by the reference
        n, A, V = self.num_vars, self.num_agents, self.max_vars_per_agent
        assign = state.variable_assignments
        B = assign.shape[0]
        acts = np.asarray(actions_array, dtype=np.int32)
        if self.action_mode == 0:                                       # env:233-244
            acts = acts.reshape(B, A)
            nv = self.action_mask.sum(axis=1).astype(np.int32)[None, :]
            no_op = acts >= nv
            safe = _jax_gather_index(np.minimum(acts, nv - 1), V)
            var = self.agent_vars[np.arange(A)[None, :], safe]
            flip = np.where(no_op, -1, var)
            counts = np.zeros((B, n + 1), dtype=np.int32)               # one_hot(...).sum(0); -1 -> slot n
            np.add.at(counts, (np.arange(B)[:, None], np.where((flip < 0) | (flip >= n), n, flip)), 1)
            new_assign = np.logical_xor(assign != 0, counts[:, :n] != 0).astype(np.int32)
        else:                                                           # env:245-250
            acts = acts.reshape(B, A, V)
            valid_vars = self.agent_vars[self.action_mask]
            valid_acts = acts[:, self.action_mask]
            new_assign = assign.copy()
            new_assign[:, valid_vars] = assign[:, valid_vars] ^ valid_acts
        status, num_unsat = self.calculate_satisfaction(new_assign, state.clauses)   # env:252
        solved = num_unsat == 0                                         # env:257
        timed_out = (state.step + 1) >= self.max_steps                  # env:258
        done = solved | timed_out
        dones = {a: done for a in self.agents}
        dones["__all__"] = done
        nxt = replace(state, variable_assignments=new_assign, clauses_satisfied_status=status,
                      num_unsatisfied=num_unsat, step=(state.step + 1).astype(np.int32),
                      done=np.repeat(done[:, None], A, axis=1))
        r = np.where(solved, np.float32(1.0), np.float32(0.0)).astype(np.float32)    # env:193
        newly = None
        if self.reward_mode == "shaped":
            # env:201-223 (commented out in the reference): PBRS + newly satisfied clauses + terminal bonus,
            # float32 with one rounding per operation (weak-typed Python floats become f32 in JAX)
            f32 = np.float32
            pot_old = (-state.num_unsatisfied).astype(f32)
            pot_new = (-num_unsat).astype(f32)
            r_pbrs = f32(self.gamma) * pot_new - pot_old                                 # env:208-210
            newly = (status & ~state.clauses_satisfied_status).astype(f32).sum(axis=1, dtype=f32)   # env:213
            r_clause_total = newly * f32(self.r_clause)                                  # env:214
            r_sat = np.where(solved, f32(self.r_sat), f32(0.0))                          # env:217
            r = ((r_pbrs + r_clause_total) + r_sat).astype(f32)                          # env:220
        rewards = {a: r for a in self.agents}
        obs = self.get_obs(nxt)
        infos = {"solved": solved, "num_unsatisfied": num_unsat,
                 "episode_step": (state.step + 1).astype(np.int32)}     # env:278-282
        if newly is not None:
            infos["newly_satisfied"] = newly.astype(np.int32)
        return obs, nxt, rewards, dones, infos

    # -- env:345-398 --------------------------------------------------------
    def get_obs_array(self, state: SATState) -> np.ndarray:
        """Observations stacked in ``env.agents`` order -> i32[B, A, D]."""
        n, m, A = self.num_vars, self.num_clauses, self.num_agents
        B = state.variable_assignments.shape[0]
        out = np.empty((B, A, 2 * n + m), dtype=np.int32)
        assign = state.variable_assignments
        sat01 = np.where(state.clauses_satisfied_status == 1, 1, 0).astype(np.int32)
        for a in range(A):
            own = np.zeros((n,), bool)
            own[self.agent_groups[self.agents[a]]] = True
            out[:, a, :n] = np.where(own[None, :], assign, -1)                                   # env:356-360
            out[:, a, n:n + m] = np.where(state.agent_clause_masks[:, a] == 1, sat01, -1)        # env:369-374
            nm = state.agent_neighbor_masks[:, a]
            out[:, a, n + m:] = np.where(nm != -1, nm * assign, -1)                              # env:382-386
        return out

    def get_obs(self, state: SATState) -> Dict[str, np.ndarray]:
        arr = self.get_obs_array(state)
        return {a: arr[:, i] for i, a in enumerate(self.agents)}


def obs_dict_to_array(env: SATEnvOracle, obs: Dict[str, np.ndarray]) -> np.ndarray:
    return np.stack([obs[a] for a in env.agents], axis=1)
