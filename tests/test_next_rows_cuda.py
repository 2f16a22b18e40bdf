"""GPU parity tests of the next-tier rows: GNN-input emitter, rollout metrics,
greedy evaluation loop -- CUDA path against the oracle restatement on the same seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import features as ofeat
from oracle.sat_env import SATEnvOracle
from tests.util import to_np

pytestmark = pytest.mark.gpu


def _setup(n, m, k, vpa, kind, B, max_steps=50, mode=0, seed=0):
    import marl_sat_b200 as M
    from marl_sat_b200.synth import mixed_ksat, uniform_ksat
    cl = uniform_ksat(B, n, m, k, seed) if kind == "uniform" else mixed_ksat(B, n, m, 3, k, seed)
    keys = np.random.default_rng(seed + 1).integers(0, 2 ** 32, size=(B, 2), dtype=np.uint64).astype(np.uint32)
    ref = SATEnvOracle(n, m, max_steps, vars_per_agent=vpa, action_mode=mode)
    env = M.SATEnv(n, m, max_steps, vars_per_agent=vpa, action_mode=mode, verbose=False)
    return M, cl, keys, ref, env


@pytest.mark.parametrize("n,m,k,vpa,kind", [(20, 91, 3, None, "uniform"), (50, 218, 3, None, "uniform"),
                                            (100, 430, 7, 7, "mixed"), (7, 12, 3, 4, "mixed")])
def test_gnn_input_matches_oracle(n, m, k, vpa, kind):
    M, cl, keys, ref, env = _setup(n, m, k, vpa, kind, B=19)
    _, st_r = ref.reset(cl, keys)
    _, st_c = env.reset(cl, keys)
    rng = np.random.default_rng(5)
    for _ in range(3):
        acts = rng.integers(0, ref.max_vars_per_agent + 1, size=(19, ref.num_agents)).astype(np.int32)
        _, st_r, *_ = ref.step_env(None, st_r, acts)
        _, st_c, *_ = env.step_env(None, st_c, acts)
    exp = ofeat.state_to_gnn_input(ref, st_r)
    got = M.gnn_input_from_state(st_c, dense_adjacency=True)
    assert np.array_equal(to_np(got.assignment), exp["assignment"])
    assert np.array_equal(to_np(got.A_pos), exp["A_pos"]) and np.array_equal(to_np(got.A_neg), exp["A_neg"])
    assert np.array_equal(to_np(got.static_var_features), exp["static_var_features"])      # bit-exact f32 division
    assert np.array_equal(to_np(got.clause_features), exp["clause_features"])
    sparse = M.gnn_input_from_state(st_c)
    assert sparse.A_pos is None and torch.equal(sparse.clause_features, got.clause_features)
    wrapper = M.SATDataWrapper(env)
    (lo, gs), ws = wrapper.reset(cl, keys)
    assert gs.clause_features.shape == (19, m, 3) and gs.static_var_features.shape == (19, n, 3)


def test_rollout_metrics_match_oracle():
    M, cl, keys, ref, env = _setup(20, 91, 3, None, "uniform", B=64, max_steps=4)
    T, B = 12, 64
    vec = M.VecSATEnv(env, torch.from_numpy(cl), B, np.array([0, 3], np.uint32), emit_obs=False)
    vec.reset()
    buf = M.RolloutBuffer(env, vec.bank, T, B)
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(T):
        acts = torch.randint(0, 5, (B, env.num_agents), generator=g, device="cuda", dtype=torch.int32)
        buf.state[t].copy_(vec.state)
        vec.step(acts, out=buf.step_outputs(t, None))
    got = M.rollout_metrics(buf.reward, buf.global_done, buf.solved, buf.num_unsatisfied, buf.episode_step)
    exp = ofeat.rollout_metrics(to_np(buf.reward), to_np(buf.global_done).astype(bool), to_np(buf.solved).astype(bool),
                                to_np(buf.num_unsatisfied), to_np(buf.episode_step))
    assert to_np(buf.global_done).sum() > 0
    for k_ in exp:
        assert abs(got[k_] - exp[k_]) <= 1e-6 * max(1.0, abs(exp[k_])), k_
    # the packed pre-step state regenerates the observation the policy saw (Transition.local_obs)
    assert buf.local_obs(3).shape == (B, env.num_agents, env.obs_dim)


def test_greedy_evaluation_loop_matches_oracle():
    M, cl, keys, ref, env = _setup(20, 40, 3, None, "uniform", B=32, max_steps=30, seed=4)   # easy: often solved

    def policy_np(obs, st):
        # deterministic "policy": every agent flips its first owned variable that sits in an unsatisfied
        # clause's neighbourhood proxy -- here simply a hash of the observation (same on both sides)
        h = (obs.astype(np.int64) * np.arange(1, obs.shape[-1] + 1)).sum(-1)
        return (h % (ref.max_vars_per_agent + 1)).astype(np.int32)

    def policy_cuda(obs, st):
        w = torch.arange(1, obs.shape[-1] + 1, device=obs.device, dtype=torch.int64)
        return ((obs.long() * w).sum(-1) % (env.max_vars_per_agent + 1)).to(torch.int32)

    ever_r, steps_r, sol_r = ofeat.evaluate_policy(policy_np, ref, cl, keys, 30)
    bank = env.make_bank(cl)
    ever_c, steps_c, sol_c = M.evaluate_policy(policy_cuda, env, bank, torch.arange(32, dtype=torch.int32), keys, 30)
    assert np.array_equal(to_np(ever_c), ever_r)
    assert np.array_equal(to_np(steps_c), steps_r)
    assert np.array_equal(to_np(sol_c), sol_r)
    assert ever_r.any()


@pytest.mark.parametrize("n,m,k,vpa,kind,tau", [(20, 91, 3, None, "uniform", 0.0), (35, 149, 3, 7, "uniform", -1.0),
                                                (30, 60, 5, 4, "dups", 0.0), (50, 218, 7, 7, "mixed", 0.0)])
def test_flip_gains_and_greedy_labels_match_brute_force(n, m, k, vpa, kind, tau):
    B = 6
    if kind == "dups":       # repeated variables, x and -x in one clause, zeros anywhere
        M, _, keys, ref, env = _setup(n, m, 3, vpa, "uniform", B)
        cl = np.random.default_rng(8).integers(-n, n + 1, size=(B, m, k)).astype(np.int32)
    else:
        M, cl, keys, ref, env = _setup(n, m, k, vpa, kind, B)
    _, st_c = env.reset(cl, keys)
    assign = to_np(st_c.variable_assignments)
    delta, labels = M.flip_gains(st_c, tau=tau)
    for b in range(B):
        exp_labels, exp_delta = ofeat.greedy_labels(ref, cl[b], assign[b], tau)
        assert np.array_equal(to_np(delta[b]), exp_delta), b
        assert np.array_equal(to_np(labels[b]), exp_labels), b


def test_fused_gnn_rollout_step_matches_separate_kernels_and_oracle():
    """msat_rollout_step_gnn: the step emits the dynamic GNN input of the state it continues from (post
    auto-reset) instead of local observations; identical to msat_gnn_dynamic on that state and to the oracle."""
    from oracle import rollout as orollout
    from oracle import threefry as otf
    M, cl, keys, ref, env = _setup(35, 149, 3, 7, "uniform", B=9, max_steps=2)
    B, P = 45, 9
    key0 = otf.prng_key(3)
    vec = M.VecSATEnv(env, cl, B, key0, emit_obs=False, compact_outputs=True, gnn_outputs=True)
    vec.reset()
    key, idx0, rk0 = orollout.initial_reset_inputs(key0, B, P)
    _, st_r = ref.reset(cl[idx0], rk0)
    rng = np.random.default_rng(1)
    for t in range(5):
        acts = rng.integers(0, ref.max_vars_per_agent + 1, size=(B, ref.num_agents)).astype(np.int32)
        ks = orollout.rollout_keys(key, B, P)
        key = ks["rng"]
        _, st_r, rew_r, done_r, _ = orollout.env_step_with_autoreset(ref, st_r, acts, cl, ks["new_problem_indices"],
                                                                      ks["reset_keys"])
        out = vec.step(torch.from_numpy(acts).cuda())
        exp = ofeat.state_to_gnn_input(ref, st_r)
        assert np.array_equal(to_np(out["gnn_assignment"]), exp["assignment"]), t
        assert np.array_equal(to_np(out["gnn_clause_features"]), exp["clause_features"]), t
        assign2, cf2 = M.dynamic_features(vec.sat_state())
        assert torch.equal(assign2, out["gnn_assignment"]) and torch.equal(cf2, out["gnn_clause_features"])
        assert np.array_equal(to_np(out["done"][:, -1]).astype(bool), done_r)
