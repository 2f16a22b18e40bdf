"""CPU tests of the oracle restatement of SATEnv / rollout / GAE: hand-derived worked examples
(SURVEY.md Appendix B), the reference's quirks (Appendix E) and invariants.  The golden fixtures
checked against the reference's own importable clause checker live in test_golden.py."""
import numpy as np
import pytest

from oracle import gae as ogae
from oracle import rollout as orollout
from oracle import threefry as tf
from oracle.sat_env import SATEnvOracle, create_agent_groups


def test_grouping_rule():
    sizes = lambda n, vpa=None: [len(v) for v in create_agent_groups(n, vpa).values()]
    assert sizes(20) == [4] * 5
    assert sizes(50) == [8] + [7] * 6                      # 4 does not divide 50 -> int(sqrt(50)) = 7 agents
    assert sizes(100) == [4] * 25
    assert sizes(250) == [17] * 10 + [16] * 5
    assert sizes(35, 7) == [7] * 5
    assert sizes(100, 7) == [7] * 10 + [6] * 5
    assert sizes(7, 4) == [4, 3]
    assert sizes(3) == [2, 1]                              # max(2, int(sqrt(3)))


def test_padding_quirk_maps():
    env = SATEnvOracle(7, 3, 10, vars_per_agent=4)
    acm, anm = env.compute_observation_maps(np.array([[[1, -2, 0], [1, -2, 3], [5, -6, 7]]], np.int32))
    assert acm[0].tolist() == [[1, 1, -1], [1, -1, 1]]     # clause 0 related to agent_1 only via -1 == -1
    assert anm[0].tolist() == [[-1] * 7, [1, 1, -1, -1, -1, -1, -1]]


def test_worked_example_end_to_end():
    env = SATEnvOracle(7, 4, 10, vars_per_agent=4)
    cl = np.array([[[1, -2, 3], [-4, 5, 0], [6, -7, 0], [-1, 4, -6]]], np.int32)
    obs, st = env.reset(cl, np.array([[0, 7]], np.uint32))
    assert st.variable_assignments[0].tolist() == [0, 1, 1, 0, 1, 0, 1]
    assert st.clauses_satisfied_status[0].tolist() == [True, True, False, True]
    assert obs["agent_0"][0].tolist() == [0, 1, 1, 0, -1, -1, -1, 1, 1, -1, 1, -1, -1, -1, -1, 1, 0, -1]
    assert obs["agent_1"][0].tolist() == [-1, -1, -1, -1, 1, 0, 1, -1, 1, 0, 1, 0, -1, -1, 0, -1, -1, -1]
    _, st, r, d, i = env.step_env(None, st, np.array([[3, 3]]))
    assert st.variable_assignments[0].tolist() == [0, 1, 1, 1, 1, 0, 1] and not d["__all__"][0]
    _, st, r, d, i = env.step_env(None, st, np.array([[4, 1]]))
    assert i["solved"][0] and d["__all__"][0] and r["agent_0"][0] == 1.0 and i["episode_step"][0] == 2


def test_timeout_uses_pre_increment_step_and_step_env_never_resets():
    env = SATEnvOracle(20, 91, max_steps=2)
    from marl_sat_b200.synth import uniform_ksat
    cl = uniform_ksat(3, 20, 91, 3, 0)
    _, st = env.reset(cl, np.zeros((3, 2), np.uint32))
    noop = np.full((3, env.num_agents), env.max_vars_per_agent, np.int32)
    _, st, _, d, i = env.step_env(None, st, noop)
    assert not d["__all__"].any() or i["solved"].any()
    _, st, _, d, i = env.step_env(None, st, noop)
    assert d["__all__"].all() and (i["episode_step"] == 2).all()
    _, st, _, d, i = env.step_env(None, st, noop)          # keeps stepping past done
    assert (st.step == 3).all() and d["__all__"].all()


def test_autoreset_select_semantics():
    from marl_sat_b200.synth import uniform_ksat
    n, m, B, P = 20, 91, 8, 5
    problems = uniform_ksat(P, n, m, 3, 1)
    env = SATEnvOracle(n, m, max_steps=1)                   # every env finishes every step
    key, idx0, rk0 = orollout.initial_reset_inputs(tf.prng_key(1), B, P)
    _, st = env.reset(problems[idx0], rk0)
    ks = orollout.rollout_keys(key, B, P)
    acts = np.zeros((B, env.num_agents), np.int32)
    fo, st2, rew, done, info = orollout.env_step_with_autoreset(env, st, acts, problems,
                                                                ks["new_problem_indices"], ks["reset_keys"])
    assert done.all() and (st2.step == 0).all() and not st2.done.any()
    assert np.array_equal(st2.clauses, problems[ks["new_problem_indices"]])
    assert (info["episode_step"] == 1).all()               # Transition keeps the pre-reset info


def test_gae_closed_form():
    T, B = 5, 3
    reward = np.zeros((T, B, 2), np.float32)
    reward[-1] = 1.0
    done = np.zeros((T, B), bool)
    value = np.zeros((T, B), np.float32)
    adv, tgt = ogae.calculate_gae(reward, done, value, np.zeros(B, np.float32), 0.9, 0.5)
    assert np.allclose(adv[:, 0], [(0.45) ** (T - 1 - t) for t in range(T)], rtol=1e-6)
    assert np.array_equal(adv, tgt)
    done[2] = True                                          # episode boundary cuts the recursion
    adv, _ = ogae.calculate_gae(reward, done, value, np.zeros(B, np.float32), 0.9, 0.5)
    assert np.allclose(adv[:3, 0], 0.0)
    norm = ogae.normalize_advantages(adv)
    assert abs(norm.mean()) < 1e-6 and abs(norm.std() - 1.0) < 1e-5
