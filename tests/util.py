"""Shared helpers for the parity tests: oracle <-> CUDA comparison through the public API."""
import numpy as np
import torch

STATE_LEAVES = ("variable_assignments", "clauses_satisfied_status", "num_unsatisfied", "step", "done", "clauses",
                "agent_clause_masks", "agent_neighbor_masks", "literal_to_agent_idx")


def to_np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def assert_state_equal(cuda_state, oracle_state, what=""):
    for leaf in STATE_LEAVES:
        got = to_np(getattr(cuda_state, leaf))
        exp = np.asarray(getattr(oracle_state, leaf))
        assert got.shape == exp.shape, f"{what}{leaf}: shape {got.shape} != {exp.shape}"
        assert np.array_equal(got.astype(exp.dtype), exp), f"{what}{leaf} differs"


def assert_obs_equal(cuda_obs_dict, oracle_obs_dict, agents, what=""):
    for a in agents:
        got, exp = to_np(cuda_obs_dict[a]), np.asarray(oracle_obs_dict[a])
        assert got.dtype == np.int32
        assert np.array_equal(got, exp), f"{what}obs[{a}] differs"
