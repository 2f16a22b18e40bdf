"""NumPy restatement of the GNN-input construction and rollout metrics (test oracle only).

``create_static_graph``: /root/reference/src/utils/graph_constructor.py:93-114;
``state_to_gnn_input``: /root/reference/src/learners/mappo_gnn_sat_learner.py:149-195;
``rollout_metrics``: learner:661-686; ``evaluate_policy``: /root/reference/src/runners/mappo_runner.py:30-73.
Pinned by reference-generated fixtures (``tests/test_golden_env.py``; see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

from .sat_env import SATEnvOracle, SATState, _jax_gather_index


def create_static_graph(num_vars: int, num_clauses: int, clauses: np.ndarray):
    """Batched over formulas: A_pos / A_neg f32[P,n,m] by scatter-add (duplicates accumulate; a 0
    literal adds 0.0 to the wrapped index -1)."""
    P, m, k = clauses.shape
    vi = _jax_gather_index(np.abs(clauses) - 1, num_vars)
    ci = np.broadcast_to(np.arange(m)[None, :, None], clauses.shape)
    pi = np.broadcast_to(np.arange(P)[:, None, None], clauses.shape)
    a_pos = np.zeros((P, num_vars, m), np.float32)
    a_neg = np.zeros((P, num_vars, m), np.float32)
    np.add.at(a_pos, (pi, vi, ci), np.where(clauses > 0, 1.0, 0.0).astype(np.float32))
    np.add.at(a_neg, (pi, vi, ci), np.where(clauses < 0, 1.0, 0.0).astype(np.float32))
    return a_pos, a_neg


def state_to_gnn_input(env: SATEnvOracle, state: SATState):
    a_pos, a_neg = create_static_graph(env.num_vars, env.num_clauses, state.clauses)
    m = np.float32(env.num_clauses)
    pos_deg = a_pos.sum(axis=2, keepdims=True) / m                         # learner:155-158
    neg_deg = a_neg.sum(axis=2, keepdims=True) / m
    svf = np.concatenate([pos_deg, neg_deg, np.zeros_like(pos_deg)], axis=-1).astype(np.float32)
    vi = _jax_gather_index(np.abs(state.clauses) - 1, env.num_vars)        # learner:179-183
    B = vi.shape[0]
    a_lit = np.take_along_axis(state.variable_assignments, vi.reshape(B, -1), axis=1).reshape(vi.shape)
    truth = ((state.clauses > 0) & (a_lit == 1)) | ((state.clauses < 0) & (a_lit == 0))
    n_true = truth.sum(axis=2)
    cf = np.stack([state.clauses_satisfied_status.astype(np.int32).astype(np.float32),
                   n_true.astype(np.float32) / np.float32(3.0),            # learner:185
                   np.ones(n_true.shape, np.float32)], axis=-1)
    return {"static_var_features": svf, "assignment": state.variable_assignments, "clause_features": cf,
            "A_pos": a_pos, "A_neg": a_neg}


def rollout_metrics(reward, done, solved, num_unsatisfied, episode_step):
    """learner:664-686 on [T,B(,A)] arrays."""
    team = reward[:, :, 0] if reward.ndim == 3 else reward
    mean_return = team.sum(axis=0).mean()
    finished = done.sum()
    solved_at_finish = solved & done
    n_solved = solved_at_finish.sum()
    return {"mean_episodic_return": float(mean_return),
            "solve_rate": float(n_solved / max(finished, 1.0)),
            "avg_unsatisfied_clauses": float((num_unsatisfied * done).sum() / max(finished, 1.0)),
            "avg_steps_to_solve": float((episode_step * solved_at_finish).sum() / max(n_solved, 1.0))}


def evaluate_policy(policy_fn, env: SATEnvOracle, clauses, keys, max_steps):
    """runner:30-73 batched over problems (no auto-reset; first solved step wins)."""
    obs, st = env.reset(clauses, keys)
    B = clauses.shape[0]
    solved_hist, assign_hist = [], []
    for _ in range(max_steps):
        acts = policy_fn(np.stack([obs[a] for a in env.agents], 1), st)
        obs, st, _, _, info = env.step_env(None, st, acts)
        solved_hist.append(info["solved"])
        assign_hist.append(st.variable_assignments)
    solved_hist, assign_hist = np.stack(solved_hist), np.stack(assign_hist)
    ever = solved_hist.any(axis=0)                                         # runner:57
    first = solved_hist.argmax(axis=0)                                     # runner:60
    solution = np.where(ever[:, None], assign_hist[first, np.arange(B)], 0)
    steps = np.where(ever, first + 1, max_steps)                           # runner:67
    return ever, steps.astype(np.int32), solution.astype(np.int32)


def greedy_labels(env: SATEnvOracle, clauses: np.ndarray, assignments: np.ndarray, tau: float):
    """``compute_joint_labels_parallel_greedy`` (behavioral_cloning.py:54-100) for ONE env, by brute force
    like the reference: flip each owned variable, recompute the unsatisfied count.  Also returns the
    per-variable deltas."""
    cl = clauses[None]
    _, base = env.calculate_satisfaction(assignments[None].astype(np.int32), cl)
    deltas = np.zeros(env.num_vars, np.int32)
    for v in range(env.num_vars):
        tmp = assignments.copy()
        tmp[v] ^= 1
        _, new = env.calculate_satisfaction(tmp[None].astype(np.int32), cl)
        deltas[v] = int(new[0]) - int(base[0])
    labels = []
    for i in range(env.num_agents):
        valid = np.flatnonzero(env.action_mask[i])
        best_delta, best = 0.0, env.max_vars_per_agent
        for j, gv in enumerate(env.agent_vars[i][valid]):
            delta = float(deltas[gv])
            if delta < best_delta:
                best_delta, best = delta, valid[j]
        labels.append(best if best_delta < tau else env.max_vars_per_agent)
    return np.array(labels, np.int32), deltas
