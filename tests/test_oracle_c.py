"""Cross-check of the two independent CPU restatements: the NumPy oracle (oracle/sat_env.py,
oracle/rollout.py, oracle/threefry.py) and the plain-C one (oracle/sat_env_c.c) must agree bit for bit on
reset, step_env, the auto-resetting rollout step and the key chain, including the reference's quirks
(literal-0 padding vs padded agents, out-of-range actions, steps past done)."""
import numpy as np
import pytest

from oracle import rollout as orollout
from oracle import threefry as otf
from oracle.c_port import SATEnvOracleC
from oracle.sat_env import SATEnvOracle


def _same_state(st_c, st_n):
    pairs = [("assign", "variable_assignments"), ("status", "clauses_satisfied_status"), ("nunsat", "num_unsatisfied"),
             ("step", "step"), ("done", "done"), ("clauses", "clauses"), ("acm", "agent_clause_masks"),
             ("anm", "agent_neighbor_masks"), ("l2a", "literal_to_agent_idx")]
    for c_name, n_name in pairs:
        a, b = st_c[c_name], np.asarray(getattr(st_n, n_name))
        assert np.array_equal(a.astype(b.dtype), b), c_name


@pytest.mark.parametrize("seed", range(12))
def test_c_and_numpy_restatements_agree(seed):
    rng = np.random.default_rng(500 + seed)
    n, m, k = int(rng.integers(1, 60)), int(rng.integers(1, 120)), int(rng.integers(1, 6))
    vpa = [None, 1, 3, 7][seed % 4]
    if vpa is not None and vpa > n:
        vpa = n
    mode = seed % 2
    B, P, max_steps = int(rng.integers(1, 12)), int(rng.integers(1, 6)), int(rng.integers(1, 4))
    problems = rng.integers(-n, n + 1, size=(P, m, k)).astype(np.int32)
    problems[rng.random((P, m)) < 0.05] = 0
    ref = SATEnvOracle(n, m, max_steps, vars_per_agent=vpa, action_mode=mode)
    cen = SATEnvOracleC(n, m, max_steps, vars_per_agent=vpa, action_mode=mode)
    assert np.array_equal(cen.agent_vars, ref.agent_vars) and cen.A == ref.num_agents
    key, idx0, rk0 = orollout.initial_reset_inputs(otf.prng_key(seed), B, P)
    obs_n, st_n = ref.reset(problems[idx0], rk0)
    st_c = cen.reset(problems[idx0], rk0)
    _same_state(st_c, st_n)
    assert np.array_equal(st_c["obs"], np.stack([obs_n[a] for a in ref.agents], 1))
    A, V = ref.num_agents, ref.max_vars_per_agent
    for t in range(5):
        acts = (rng.integers(-2, V + 2, size=(B, A)) if mode == 0 else rng.integers(0, 2, size=(B, A, V))).astype(np.int32)
        ks = orollout.rollout_keys(key, B, P)
        chain, idx_c, keys_c = cen.rollout_keys(key, B, P)
        assert np.array_equal(chain, np.concatenate([ks["rng"], ks["act_key"], ks["step_key"], ks["prob_key"], ks["reset_key"]]))
        assert np.array_equal(idx_c, ks["new_problem_indices"]) and np.array_equal(keys_c, ks["reset_keys"])
        key = ks["rng"]
        if t % 2 == 0:      # auto-resetting rollout step
            fo, st_n, rew, done, info = orollout.env_step_with_autoreset(ref, st_n, acts, problems, idx_c, keys_c)
            out = cen.step(st_c, acts, problems, idx_c, keys_c)
        else:               # plain step_env (keeps stepping past done)
            obs_n, st_n, rewards, dones, info = ref.step_env(None, st_n, acts)
            fo = np.stack([obs_n[a] for a in ref.agents], 1)
            rew = np.stack([rewards[a] for a in ref.agents], -1)
            done = dones["__all__"]
            out = cen.step(st_c, acts)
        _same_state(st_c, st_n)
        assert np.array_equal(st_c["obs"], fo)
        assert np.array_equal(out["reward"], rew) and np.array_equal(out["done_all"].astype(bool), done)
        assert np.array_equal(out["solved"].astype(bool), info["solved"])
        assert np.array_equal(out["num_unsatisfied"], info["num_unsatisfied"])
        assert np.array_equal(out["episode_step"], info["episode_step"])
