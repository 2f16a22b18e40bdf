"""Parity against fixtures produced by EXECUTING THE UNMODIFIED REFERENCE SOURCE
(``tests/golden/make_golden_env.py``: the reference's ``SATEnv``, ``SATDataWrapper``, ``_env_step``,
``_calculate_gae``, normalisation, metric block, ``evaluate_policy`` and the BC labeller, run on the NumPy
stand-ins of ``tests/ref_shim``).  CPU tests: the oracle restatement reproduces the fixtures.  GPU tests
(``-m gpu``): the CUDA path, through the C ABI, reproduces the same fixtures.

Bit-exact for every integer / bool / index leaf and for the f32 feature divisions; GAE within 1e-5
relative (the tolerance BASELINE.json's north_star states).
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import features as ofeat
from oracle import gae as ogae
from oracle import rollout as orollout
from oracle.sat_env import SATEnvOracle

GOLD = Path(__file__).resolve().parent / "golden"
ROLLOUT_CASES = ["c1_uf20_mode0", "c1_uf20_mode1", "loose12_mode0", "loose12_mode1", "yaml_uf35_vpa7", "c2_uf50",
                 "c3_uf100", "c4_uf250", "c5_mixedk_vpa7", "c5_mixedk_mode1", "wild_actions_uneven", "pad_quirk_n7",
                 "single_agent", "one_var_agents", "shaped_loose12", "shaped_uf50_mode1"]
STEPPING_CASES = ["past_done_mode0", "past_done_mode1"]
EVAL_CASES = ["eval_bc_loose", "eval_bc_uneven"]
GAE_RTOL = 1e-5


def load(name):
    fx = dict(np.load(GOLD / f"env_{name}.npz"))
    n, m, k, P, B, T, max_steps, vpa, mode, A, V = (int(x) for x in fx["meta"])
    fx["cfg"] = dict(n=n, m=m, k=k, P=P, B=B, T=T, max_steps=max_steps, vpa=None if vpa < 0 else vpa, mode=mode, A=A, V=V)
    sh = fx.get("shaped")
    # the reference's commented-out shaped reward (env:201-223), run uncommented by the fixture generator
    fx["reward_kw"] = ({} if sh is None or sh[0] == 0 else
                       dict(r_clause=float(sh[1]), r_sat=float(sh[2]), gamma=float(sh[3]), reward_mode="shaped"))
    return fx


def eq(got, exp, what):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, f"{what}: shape {got.shape} != {exp.shape}"
    if exp.dtype == np.bool_:
        got = got.astype(bool)
    assert np.array_equal(got, exp.astype(got.dtype) if got.dtype != exp.dtype else exp), f"{what} differs"


def close(got, exp, what, rtol=GAE_RTOL):
    got, exp = np.asarray(got, np.float64), np.asarray(exp, np.float64)
    scale = max(1.0, float(np.abs(exp).max()))
    assert np.allclose(got, exp, rtol=rtol, atol=rtol * scale), f"{what}: max err {np.abs(got - exp).max()}"


def test_fixtures_are_reference_outputs():
    """Sanity of the fixtures themselves: every BASELINE config shape is present, episodes end and
    restart inside the rollouts, some are solved, and the F6 padding quirk is exercised."""
    total_resets = total_solved = 0
    assert np.unique(load("shaped_loose12")["tr_reward"]).size > 6      # the shaped reward is not 0/1
    for name in ROLLOUT_CASES:
        fx = load(name)
        assert fx["tr_local_obs"].dtype == np.int32 and fx["tr_reward"].dtype == np.float32
        assert fx["advantages"].dtype == np.float32
        total_resets += int(fx["tr_global_done"].sum())
        total_solved += int(fx["tr_info_solved"].sum())
    assert total_resets > 200 and total_solved > 40
    fx = load("c5_mixedk_vpa7")
    # F6: a clause whose only link to a short agent is the 0 padding (-1 == -1) is still "related"
    cl, av = fx["clauses"][fx["initial_indices"]], fx["agent_vars"]
    short = np.flatnonzero((av == -1).any(axis=1))
    padded = (cl == 0).any(axis=2)
    assert short.size and padded.any()
    acm = fx["s0_agent_clause_masks"]
    for b in range(cl.shape[0]):
        for a in short:
            assert (acm[b, a][padded[b]] == 1).all()


# ---------------------------------------------------------------------------------------------------------
# oracle == reference fixtures (CPU)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_oracle_rollout_matches_reference(name):
    fx = load(name)
    c = fx["cfg"]
    env = SATEnvOracle(c["n"], c["m"], c["max_steps"], vars_per_agent=c["vpa"], action_mode=c["mode"], **fx["reward_kw"])
    eq(env.agent_vars, fx["agent_vars"], "agent_vars")
    eq(env.action_mask, fx["action_mask"], "action_mask")
    eq(env.variable_to_agent_idx, fx["variable_to_agent_idx"], "variable_to_agent_idx")
    tr, fin = orollout.rollout_T(env, fx["clauses"], fx["key0"], fx["actions"], fx["values"][:-1])
    ini = fin["initial"]
    eq(ini["problem_idx"], fx["initial_indices"], "initial_indices")
    eq(ini["reset_keys"], fx["initial_reset_keys"], "initial_reset_keys")
    eq(ini["rng"], fx["rng_after_init"], "rng after the initial split")
    eq(ini["obs"], fx["obs0"], "obs0")
    for leaf in ("variable_assignments", "clauses_satisfied_status", "num_unsatisfied", "agent_clause_masks",
                 "agent_neighbor_masks", "literal_to_agent_idx", "step", "done"):
        eq(getattr(ini["state"], leaf), fx[f"s0_{leaf}"], f"s0.{leaf}")
    gs0 = ofeat.state_to_gnn_input(env, ini["state"])
    eq(gs0["static_var_features"], fx["gs0_static_var_features"], "static_var_features")
    eq(gs0["clause_features"], fx["gs0_clause_features"], "clause_features(reset)")
    if "graph0_A_pos" in fx:
        a_pos, a_neg = ofeat.create_static_graph(c["n"], c["m"], fx["clauses"][:1])
        eq(a_pos[0], fx["graph0_A_pos"], "A_pos")
        eq(a_neg[0], fx["graph0_A_neg"], "A_neg")
    # Transition (learner:467-478)
    eq(tr["act_key"], fx["act_keys"], "act_key chain")
    eq(tr["local_obs"], fx["tr_local_obs"], "Transition.local_obs (pre-step)")
    eq(tr["gs_assignment"], fx["tr_gs_assignment"], "Transition.global_state.assignment")
    eq(tr["gs_clause_features"], fx["tr_gs_clause_features"], "Transition.global_state.clause_features")
    eq(tr["reward"], fx["tr_reward"], "Transition.reward")
    eq(tr["global_done"], fx["tr_global_done"], "Transition.global_done")
    eq(tr["info_solved"], fx["tr_info_solved"], "info.solved")
    eq(tr["info_num_unsatisfied"], fx["tr_info_num_unsatisfied"], "info.num_unsatisfied")
    eq(tr["info_episode_step"], fx["tr_info_episode_step"], "info.episode_step")
    # final carry
    eq(fin["obs"], fx["final_obs"], "final obs")
    eq(fin["rng"], fx["final_rng"], "final rng")
    for leaf in ("variable_assignments", "clauses_satisfied_status", "num_unsatisfied", "step", "done", "clauses",
                 "agent_clause_masks", "agent_neighbor_masks", "literal_to_agent_idx"):
        eq(getattr(fin["state"], leaf), fx[f"final_{leaf}"], f"final.{leaf}")
    gsf = ofeat.state_to_gnn_input(env, fin["state"])
    eq(gsf["clause_features"], fx["final_gs_clause_features"], "final clause_features")
    eq(gsf["static_var_features"], fx["final_gs_static_var_features"], "final static_var_features")
    # GAE / normalisation / metrics
    g, lam = fx["gamma_lambda"]
    adv, tgt = ogae.calculate_gae(tr["reward"], tr["global_done"], fx["tr_value"], fx["last_val"], g, lam)
    eq(adv, fx["advantages"], "advantages (same f32 operation order -> bit-exact on CPU)")
    eq(tgt, fx["targets"], "targets")
    close(ogae.normalize_advantages(adv), fx["advantages_normalized"], "normalised advantages")
    met = ofeat.rollout_metrics(tr["reward"], tr["global_done"], tr["info_solved"], tr["info_num_unsatisfied"],
                                tr["info_episode_step"])
    for k_ in ("mean_episodic_return", "solve_rate", "avg_unsatisfied_clauses", "avg_steps_to_solve"):
        close(met[k_], fx[f"metric_{k_}"], k_, rtol=1e-6)


@pytest.mark.parametrize("name", STEPPING_CASES)
def test_oracle_stepping_past_done_matches_reference(name):
    fx = load(name)
    c = fx["cfg"]
    env = SATEnvOracle(c["n"], c["m"], c["max_steps"], vars_per_agent=c["vpa"], action_mode=c["mode"])
    obs, st = env.reset(fx["clauses"], fx["keys"])
    for t in range(c["T"]):
        obs, st, rewards, dones, infos = env.step_env(None, st, fx["actions"][t])
        eq(np.stack([obs[a] for a in env.agents], 1), fx["obs"][t], f"obs[{t}]")
        eq(st.variable_assignments, fx["assign"][t], f"assign[{t}]")
        eq(st.clauses_satisfied_status, fx["status"][t], f"status[{t}]")
        eq(st.num_unsatisfied, fx["nunsat"][t], f"nunsat[{t}]")
        eq(st.step, fx["step"][t], f"step[{t}]")
        eq(st.done, fx["done"][t], f"done[{t}]")
        eq(np.stack([rewards[a] for a in env.agents], -1), fx["reward"][t], f"reward[{t}]")
        eq(dones["__all__"], fx["done_all"][t], f"done_all[{t}]")
        eq(infos["solved"], fx["solved"][t], f"solved[{t}]")
        eq(infos["episode_step"], fx["episode_step"][t], f"episode_step[{t}]")


@pytest.mark.parametrize("name", EVAL_CASES)
def test_oracle_eval_and_bc_labels_match_reference(name):
    fx = load(name)
    c = fx["cfg"]
    env = SATEnvOracle(c["n"], c["m"], c["max_steps"], vars_per_agent=c["vpa"])
    P = c["P"]
    # runner:31: key, reset_key = split(key); the reset uses reset_key
    from oracle import threefry
    reset_keys = np.stack([threefry.split(fx["keys"][p])[1] for p in range(P)])
    step = {"t": 0}

    def policy(obs, st):
        a = fx["logits"][:, step["t"]].argmax(axis=-1).astype(np.int32)
        step["t"] += 1
        return a
    ever, steps, sol = ofeat.evaluate_policy(policy, env, fx["clauses"], reset_keys, c["max_steps"])
    eq(ever, fx["ever_solved"], "was_ever_solved")
    eq(steps, fx["steps_to_solve"], "steps_to_solve")
    eq(sol, fx["solution"], "solution_assignments")
    for ti, tau in enumerate(fx["bc_taus"]):
        for p in range(P):
            lab, _ = ofeat.greedy_labels(env, fx["clauses"][p], fx["bc_assignments"][p], float(tau))
            eq(lab, fx["bc_labels"][ti, p], f"bc labels tau={tau} p={p}")


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_c_port_rollout_matches_reference(name):
    """The plain-C/OpenMP restatement (the compiled CPU baseline of bench.py) against the same fixtures."""
    from oracle.c_port import SATEnvOracleC
    fx = load(name)
    c = fx["cfg"]
    if fx["reward_kw"]:
        pytest.skip("the C port implements the active (sparse) reward only")
    cen = SATEnvOracleC(c["n"], c["m"], c["max_steps"], vars_per_agent=c["vpa"], action_mode=c["mode"])
    st = cen.reset(fx["clauses"][fx["initial_indices"]], fx["initial_reset_keys"])
    eq(st["obs"], fx["obs0"], "obs0")
    rng = fx["rng_after_init"]
    for t in range(c["T"]):
        eq(st["obs"], fx["tr_local_obs"][t], f"local_obs[{t}]")
        chain, idx, keys = cen.rollout_keys(rng, c["B"], c["P"])
        rng = chain[0:2].copy()
        eq(chain[2:4], fx["act_keys"][t], f"act_key[{t}]")
        out = cen.step(st, fx["actions"][t], fx["clauses"], idx, keys)
        eq(out["reward"], fx["tr_reward"][t], f"reward[{t}]")
        eq(out["done_all"], fx["tr_global_done"][t], f"done[{t}]")
        eq(out["solved"], fx["tr_info_solved"][t], f"solved[{t}]")
        eq(out["num_unsatisfied"], fx["tr_info_num_unsatisfied"][t], f"num_unsatisfied[{t}]")
        eq(out["episode_step"], fx["tr_info_episode_step"][t], f"episode_step[{t}]")
    eq(st["obs"], fx["final_obs"], "final obs")
    eq(rng, fx["final_rng"], "final rng")
    for c_name, leaf in [("assign", "variable_assignments"), ("status", "clauses_satisfied_status"),
                         ("nunsat", "num_unsatisfied"), ("step", "step"), ("done", "done"), ("clauses", "clauses"),
                         ("acm", "agent_clause_masks"), ("anm", "agent_neighbor_masks"), ("l2a", "literal_to_agent_idx")]:
        eq(st[c_name], fx[f"final_{leaf}"], f"final.{leaf}")


# ---------------------------------------------------------------------------------------------------------
# CUDA path == reference fixtures (GPU, through the C ABI)
# ---------------------------------------------------------------------------------------------------------
def _to_np(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("fused_keys", [True, False])
@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_cuda_rollout_matches_reference(name, fused_keys):
    import torch
    import marl_sat_b200 as M
    from marl_sat_b200 import features as F
    from tests.util import STATE_LEAVES
    fx = load(name)
    c = fx["cfg"]
    T, B = c["T"], c["B"]
    env = M.SATEnv(c["n"], c["m"], c["max_steps"], vars_per_agent=c["vpa"], action_mode=c["mode"], verbose=False,
                   **fx["reward_kw"])
    eq(_to_np(env.agent_vars), fx["agent_vars"], "agent_vars")
    eq(_to_np(env.action_mask), fx["action_mask"], "action_mask")
    eq(_to_np(env.variable_to_agent_idx), fx["variable_to_agent_idx"], "variable_to_agent_idx")
    vec = M.VecSATEnv(env, torch.from_numpy(fx["clauses"]), B, fx["key0"], fused_keys=fused_keys)
    obs0 = vec.reset()
    eq(_to_np(obs0), fx["obs0"], "obs0")
    eq(M.env.u32_to_numpy(vec.keys.rng), fx["rng_after_init"], "rng after the initial split")
    eq(_to_np(vec.new_problem_idx), fx["initial_indices"], "initial_indices")
    eq(M.env.u32_to_numpy(vec.reset_keys), fx["initial_reset_keys"], "initial_reset_keys")
    s0 = vec.sat_state()
    for leaf in ("variable_assignments", "clauses_satisfied_status", "num_unsatisfied", "agent_clause_masks",
                 "agent_neighbor_masks", "literal_to_agent_idx", "step", "done"):
        eq(_to_np(getattr(s0, leaf)), fx[f"s0_{leaf}"], f"s0.{leaf}")
    gs0 = M.gnn_input_from_state(s0, dense_adjacency="graph0_A_pos" in fx)
    eq(_to_np(gs0.static_var_features), fx["gs0_static_var_features"], "static_var_features")
    eq(_to_np(gs0.clause_features), fx["gs0_clause_features"], "clause_features(reset)")
    if "graph0_A_pos" in fx:
        sg = F.static_graph(vec.bank, dense_adjacency=True)
        eq(_to_np(sg.A_pos[0]), fx["graph0_A_pos"], "A_pos")
        eq(_to_np(sg.A_neg[0]), fx["graph0_A_neg"], "A_neg")

    buf = M.RolloutBuffer(env, vec.bank, T, B)
    actions = torch.from_numpy(fx["actions"]).cuda()
    values = torch.from_numpy(fx["values"]).cuda()
    act_keys = []

    def policy(t, v):
        # the reference draws act_key BEFORE the step (learner:397); the fused step advances the chain,
        # so the key is read back after the step below
        return actions[t], values[t], None

    # step by step (instead of buf.collect) to also read the per-step act_key
    for t in range(T):
        buf.state[t].copy_(vec.state)
        buf.action[t].copy_(actions[t])
        buf.value[t].copy_(values[t])
        vec.step(buf.action[t], out=buf.step_outputs(t, vec.out["obs"]))
        act_keys.append(M.env.u32_to_numpy(vec.keys.act_key).copy())
    eq(np.stack(act_keys), fx["act_keys"], "act_key chain")
    for t in range(T):
        eq(_to_np(buf.local_obs(t)), fx["tr_local_obs"][t], f"Transition.local_obs[{t}] (pre-step)")
        a, cf = F.dynamic_features(buf.pre_step_state(t))
        eq(_to_np(a), fx["tr_gs_assignment"][t], f"Transition.global_state.assignment[{t}]")
        eq(_to_np(cf), fx["tr_gs_clause_features"][t], f"Transition.global_state.clause_features[{t}]")
    eq(_to_np(buf.reward_per_agent), fx["tr_reward"], "Transition.reward")
    eq(_to_np(buf.global_done), fx["tr_global_done"], "Transition.global_done")
    eq(_to_np(buf.info["solved"]), fx["tr_info_solved"], "info.solved")
    eq(_to_np(buf.info["num_unsatisfied"]), fx["tr_info_num_unsatisfied"], "info.num_unsatisfied")
    eq(_to_np(buf.info["episode_step"]), fx["tr_info_episode_step"], "info.episode_step")
    eq(_to_np(buf.action), fx["tr_action"], "Transition.action")
    # final carry
    eq(_to_np(vec.out["obs"]), fx["final_obs"], "final obs")
    eq(M.env.u32_to_numpy(vec.keys.rng), fx["final_rng"], "final rng")
    fs = vec.sat_state()
    for leaf in STATE_LEAVES:
        eq(_to_np(getattr(fs, leaf)), fx[f"final_{leaf}"], f"final.{leaf}")
    gsf = M.gnn_input_from_state(fs)
    eq(_to_np(gsf.assignment), fx["final_gs_assignment"], "final assignment")
    eq(_to_np(gsf.clause_features), fx["final_gs_clause_features"], "final clause_features")
    eq(_to_np(gsf.static_var_features), fx["final_gs_static_var_features"], "final static_var_features")
    # GAE from the buffer (learner:504-532)
    g, lam = fx["gamma_lambda"]
    stats = torch.zeros(3, dtype=torch.float64, device="cuda")
    adv, tgt = M.calculate_gae(buf.reward_per_agent, buf.global_done, buf.value, values[T], g, lam, stats=stats)
    close(_to_np(adv), fx["advantages"], "advantages")
    close(_to_np(tgt), fx["targets"], "targets")
    close(_to_np(M.normalize_advantages(adv.clone(), stats=stats)), fx["advantages_normalized"], "normalised advantages",
          rtol=2e-5)
    met = M.rollout_metrics(buf.reward, buf.global_done, buf.solved, buf.num_unsatisfied, buf.episode_step)
    for k_ in ("mean_episodic_return", "solve_rate", "avg_unsatisfied_clauses", "avg_steps_to_solve"):
        close(met[k_], fx[f"metric_{k_}"], k_, rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ROLLOUT_CASES[:6])
def test_cuda_collect_equals_stepwise(name):
    """``RolloutBuffer.collect`` (the product's rollout loop) fills the same Transition as the fixtures."""
    import torch
    import marl_sat_b200 as M
    fx = load(name)
    c = fx["cfg"]
    env = M.SATEnv(c["n"], c["m"], c["max_steps"], vars_per_agent=c["vpa"], action_mode=c["mode"], verbose=False)
    vec = M.VecSATEnv(env, torch.from_numpy(fx["clauses"]), c["B"], fx["key0"])
    vec.reset()
    buf = M.RolloutBuffer(env, vec.bank, c["T"], c["B"])
    actions = torch.from_numpy(fx["actions"]).cuda()
    values = torch.from_numpy(fx["values"]).cuda()
    buf.collect(vec, lambda t, v: (actions[t], values[t], None))
    eq(_to_np(buf.reward_per_agent), fx["tr_reward"], "Transition.reward")
    eq(_to_np(buf.global_done), fx["tr_global_done"], "Transition.global_done")
    eq(_to_np(buf.value), fx["tr_value"], "Transition.value")
    eq(_to_np(buf.info["episode_step"]), fx["tr_info_episode_step"], "info.episode_step")
    for t in (0, c["T"] // 2, c["T"] - 1):
        eq(_to_np(buf.local_obs(t)), fx["tr_local_obs"][t], f"local_obs[{t}]")
    eq(_to_np(vec.out["obs"]), fx["final_obs"], "final obs")


@pytest.mark.gpu
@pytest.mark.parametrize("name", STEPPING_CASES)
def test_cuda_stepping_past_done_matches_reference(name):
    import marl_sat_b200 as M
    fx = load(name)
    c = fx["cfg"]
    env = M.SATEnv(c["n"], c["m"], c["max_steps"], vars_per_agent=c["vpa"], action_mode=c["mode"], verbose=False)
    obs, st = env.reset(fx["clauses"], fx["keys"])
    for t in range(c["T"]):
        obs, st, rewards, dones, infos = env.step_env(None, st, fx["actions"][t])
        eq(np.stack([_to_np(obs[a]) for a in env.agents], 1), fx["obs"][t], f"obs[{t}]")
        eq(_to_np(st.variable_assignments), fx["assign"][t], f"assign[{t}]")
        eq(_to_np(st.clauses_satisfied_status), fx["status"][t], f"status[{t}]")
        eq(_to_np(st.num_unsatisfied), fx["nunsat"][t], f"nunsat[{t}]")
        eq(_to_np(st.step), fx["step"][t], f"step[{t}]")
        eq(_to_np(st.done), fx["done"][t], f"done[{t}]")
        eq(np.stack([_to_np(rewards[a]) for a in env.agents], -1), fx["reward"][t], f"reward[{t}]")
        eq(_to_np(dones["__all__"]), fx["done_all"][t], f"done_all[{t}]")
        eq(_to_np(infos["solved"]), fx["solved"][t], f"solved[{t}]")
        eq(_to_np(infos["num_unsatisfied"]), fx["info_nunsat"][t], f"info nunsat[{t}]")
        eq(_to_np(infos["episode_step"]), fx["episode_step"][t], f"episode_step[{t}]")


@pytest.mark.gpu
@pytest.mark.parametrize("name", EVAL_CASES)
def test_cuda_eval_and_bc_labels_match_reference(name):
    import torch
    import marl_sat_b200 as M
    from oracle import threefry
    fx = load(name)
    c = fx["cfg"]
    P = c["P"]
    env = M.SATEnv(c["n"], c["m"], c["max_steps"], vars_per_agent=c["vpa"], verbose=False)
    bank = env.make_bank(fx["clauses"])
    reset_keys = np.stack([threefry.split(fx["keys"][p])[1] for p in range(P)])      # runner:31
    logits = torch.from_numpy(fx["logits"]).cuda()
    step = {"t": 0}

    def policy(obs, st):
        a = logits[:, step["t"]].argmax(dim=-1).to(torch.int32)                       # runner:41
        step["t"] += 1
        return a
    ever, steps, sol = M.evaluate_policy(policy, env, bank, torch.arange(P, dtype=torch.int32, device="cuda"),
                                         reset_keys, c["max_steps"])
    eq(_to_np(ever), fx["ever_solved"], "was_ever_solved")
    eq(_to_np(steps), fx["steps_to_solve"], "steps_to_solve")
    eq(_to_np(sol), fx["solution"], "solution_assignments")
    # BC expert labels (behavioral_cloning.py:54-100) from a state whose assignment is the fixture's
    _, st = env.reset_from_bank(bank, torch.arange(P, dtype=torch.int32, device="cuda"), reset_keys)
    aw = (c["n"] + 31) // 32
    bits = np.zeros((P, aw), np.uint32)
    for p in range(P):
        for v in np.flatnonzero(fx["bc_assignments"][p]):
            bits[p, v >> 5] |= np.uint32(1) << np.uint32(v & 31)
    st.packed[:, :aw] = torch.from_numpy(bits.view(np.int32)).cuda()
    for ti, tau in enumerate(fx["bc_taus"]):
        _, labels = M.flip_gains(M.SATState(env, bank, st.packed, True), tau=float(tau))
        eq(_to_np(labels), fx["bc_labels"][ti], f"bc labels tau={tau}")
