"""GPU parity tests: the CUDA path (through the C ABI / the drop-in SATEnv) against the NumPy oracle on
the same seeded inputs.  Bit-exact for every integer/byte/index output and for the 0/1 rewards; GAE
within 1e-5 relative (the tolerance BASELINE.json's north_star states).

Run on the B200 box: ``python -m pytest tests -m gpu -x -q``.
"""
import numpy as np
import pytest
import torch

from oracle import gae as ogae
from oracle import rollout as orollout
from oracle import threefry as otf
from oracle.sat_env import SATEnvOracle
from tests.util import assert_obs_equal, assert_state_equal, to_np

pytestmark = pytest.mark.gpu


def _msat():
    import marl_sat_b200 as M
    return M


def _keys(B, seed):
    return np.random.default_rng(seed).integers(0, 2 ** 32, size=(B, 2), dtype=np.uint64).astype(np.uint32)


# (name, n, m, k, vars_per_agent, kind)
SHAPES = [
    ("uf20-91 auto (C1)", 20, 91, 3, None, "uniform"),
    ("uf35-149 vpa7 (YAML)", 35, 149, 3, 7, "uniform"),
    ("uf50-218 auto (C2)", 50, 218, 3, None, "uniform"),
    ("uf100-430 auto (C3)", 100, 430, 3, None, "uniform"),
    ("uf250-1065 auto (C4)", 250, 1065, 3, None, "uniform"),
    ("mixed k3-7 vpa7 (C5, F6 quirk)", 100, 430, 7, 7, "mixed"),
    ("tiny padded uneven", 7, 4, 3, 4, "mixed23"),
    ("one var per agent", 12, 30, 3, 1, "uniform"),
    ("single agent", 9, 20, 3, 9, "uniform"),
]


def _formulas(kind, P, n, m, k, seed):
    from marl_sat_b200.synth import mixed_ksat, uniform_ksat
    if kind == "uniform":
        return uniform_ksat(P, n, m, k, seed)
    if kind == "mixed":
        return mixed_ksat(P, n, m, 3, k, seed)
    return mixed_ksat(P, n, m, 2, k, seed)


def _random_actions(rng, env, B):
    A, V = env.num_agents, env.max_vars_per_agent
    if env.action_mode == 0:
        # in-space actions 0..V (V and anything >= group size is a no-op) plus a few beyond V
        return rng.integers(0, V + 2, size=(B, A)).astype(np.int32)
    return rng.integers(0, 2, size=(B, A, V)).astype(np.int32)


@pytest.mark.parametrize("name,n,m,k,vpa,kind", SHAPES, ids=[s[0] for s in SHAPES])
@pytest.mark.parametrize("mode", [0, 1])
def test_reset_and_step_match_oracle(name, n, m, k, vpa, kind, mode):
    M = _msat()
    B = 37 if n < 200 else 11          # odd counts exercise misaligned per-env observation blocks
    clauses = _formulas(kind, B, n, m, k, seed=n * 7 + m)
    keys = _keys(B, seed=n + 1)
    ref = SATEnvOracle(n, m, max_steps=4, vars_per_agent=vpa, action_mode=mode)
    env = M.SATEnv(n, m, max_steps=4, vars_per_agent=vpa, action_mode=mode, verbose=False)
    assert env.agents == ref.agents and env.max_vars_per_agent == ref.max_vars_per_agent
    assert np.array_equal(to_np(env.agent_vars), ref.agent_vars)
    assert np.array_equal(to_np(env.action_mask), ref.action_mask)
    assert np.array_equal(to_np(env.variable_to_agent_idx), ref.variable_to_agent_idx)

    obs_r, st_r = ref.reset(clauses, keys)
    obs_c, st_c = env.reset(torch.from_numpy(clauses), keys)
    assert_obs_equal(obs_c, obs_r, env.agents, "reset ")
    assert_state_equal(st_c, st_r, "reset ")
    assert_obs_equal(env.get_obs(st_c), obs_r, env.agents, "get_obs ")

    rng = np.random.default_rng(n + m + mode)
    for t in range(6):                 # runs past max_steps: step_env never resets (env:225-284)
        acts = _random_actions(rng, ref, B)
        obs_r, st_r, rew_r, done_r, info_r = ref.step_env(None, st_r, acts)
        obs_c, st_c, rew_c, done_c, info_c = env.step_env(None, st_c, torch.from_numpy(acts).cuda())
        assert_obs_equal(obs_c, obs_r, env.agents, f"step {t} ")
        assert_state_equal(st_c, st_r, f"step {t} ")
        for a in env.agents:
            assert to_np(rew_c[a]).dtype == np.float32
            assert np.array_equal(to_np(rew_c[a]), rew_r[a])
            assert np.array_equal(to_np(done_c[a]), done_r[a])
        assert np.array_equal(to_np(done_c["__all__"]), done_r["__all__"])
        for key in ("solved", "num_unsatisfied", "episode_step"):
            assert np.array_equal(to_np(info_c[key]), info_r[key]), key


def test_action_dict_and_unbatched_calls():
    M = _msat()
    n, m = 20, 91
    clauses = _formulas("uniform", 1, n, m, 3, 5)[0]
    key = np.array([0, 7], np.uint32)
    ref = SATEnvOracle(n, m, 10)
    env = M.SATEnv(n, m, 10, verbose=False)
    obs_r, st_r = ref.reset(clauses[None], key[None])
    obs_c, st_c = env.reset(clauses, key)                       # unbatched, like the reference signature
    for a in env.agents:
        assert tuple(obs_c[a].shape) == (env.obs_dim,)
        assert np.array_equal(to_np(obs_c[a]), obs_r[a][0])
    assert tuple(st_c.variable_assignments.shape) == (n,)
    acts = {a: torch.tensor(i % 5, dtype=torch.int32) for i, a in enumerate(env.agents)}
    wrapper = M.SATDataWrapper(env, emit_global_state=False)
    (lo, _), ws, rew, done, info = wrapper.step(None, M.GNNWrapperState(st_c, st_c.bank), acts)
    o2, s2, r2, d2, i2 = ref.step_env(None, st_r, np.array([[i % 5 for i in range(env.num_agents)]], np.int32))
    for a in env.agents:
        assert np.array_equal(to_np(lo[a]), o2[a][0])
    assert bool(done["__all__"]) == bool(d2["__all__"][0])
    assert int(info["episode_step"]) == 1
    with pytest.raises(TypeError):
        env.step(None, st_c, acts)


def test_known_worked_example():
    """SURVEY.md Appendix B second worked example (hand-derived from env:158-284)."""
    M = _msat()
    env = M.SATEnv(7, 4, 10, vars_per_agent=4, verbose=False)
    cl = np.array([[1, -2, 3], [-4, 5, 0], [6, -7, 0], [-1, 4, -6]], np.int32)
    obs, st = env.reset(cl, np.array([0, 7], np.uint32))
    assert to_np(st.variable_assignments).tolist() == [0, 1, 1, 0, 1, 0, 1]
    assert to_np(st.clauses_satisfied_status).tolist() == [True, True, False, True]
    assert int(st.num_unsatisfied) == 1
    assert to_np(st.agent_clause_masks).tolist() == [[1, 1, -1, 1], [-1, 1, 1, 1]]
    assert to_np(st.agent_neighbor_masks).tolist() == [[-1, -1, -1, -1, 1, 1, -1], [1, -1, -1, 1, -1, -1, -1]]
    assert to_np(obs["agent_0"]).tolist() == [0, 1, 1, 0, -1, -1, -1, 1, 1, -1, 1, -1, -1, -1, -1, 1, 0, -1]
    assert to_np(obs["agent_1"]).tolist() == [-1, -1, -1, -1, 1, 0, 1, -1, 1, 0, 1, 0, -1, -1, 0, -1, -1, -1]
    obs, st, rew, done, info = env.step_env(None, st, np.array([3, 3], np.int32))
    assert to_np(st.variable_assignments).tolist() == [0, 1, 1, 1, 1, 0, 1] and not bool(done["__all__"])
    obs, st, rew, done, info = env.step_env(None, st, np.array([4, 1], np.int32))
    assert to_np(st.variable_assignments).tolist() == [0, 1, 1, 1, 1, 1, 1]
    assert bool(info["solved"]) and bool(done["__all__"]) and float(rew["agent_1"]) == 1.0
    assert int(info["episode_step"]) == 2 and int(st.step) == 2


def test_negative_and_out_of_range_actions_follow_jax_indexing():
    M = _msat()
    n, m = 50, 218                      # uneven groups [8,7,7,...]: V=8, agents 1.. have a -1 pad slot
    B = 16
    clauses = _formulas("uniform", B, n, m, 3, 3)
    keys = _keys(B, 9)
    ref = SATEnvOracle(n, m, 100)
    env = M.SATEnv(n, m, 100, verbose=False)
    _, st_r = ref.reset(clauses, keys)
    _, st_c = env.reset(clauses, keys)
    rng = np.random.default_rng(0)
    acts = rng.integers(-12, 12, size=(B, ref.num_agents)).astype(np.int32)
    _, st_r, *_ = ref.step_env(None, st_r, acts)
    _, st_c, *_ = env.step_env(None, st_c, acts)
    assert_state_equal(st_c, st_r)


@pytest.mark.parametrize("fused", [False, True], ids=["separate-key-kernels", "fused-keys"])
@pytest.mark.parametrize("n,m,vpa,B,P,max_steps", [(20, 91, None, 64, 10, 3), (35, 149, 7, 33, 5, 2),
                                                   (100, 430, None, 48, 48, 2)])
def test_rollout_autoreset_and_rng_chain(n, m, vpa, B, P, max_steps, fused):
    """VecSATEnv (fused auto-reset + device RNG chain) against the oracle restatement of learner:383-480."""
    M = _msat()
    problems = _formulas("uniform", P, n, m, 3, seed=11)
    ref = SATEnvOracle(n, m, max_steps, vars_per_agent=vpa)
    env = M.SATEnv(n, m, max_steps, vars_per_agent=vpa, verbose=False)
    key0 = otf.prng_key(42)
    vec = M.VecSATEnv(env, torch.from_numpy(problems), B, key0, fused_keys=fused)
    obs_c = vec.reset()
    key, idx0, rk0 = orollout.initial_reset_inputs(key0, B, P)
    obs_r, st_r = ref.reset(problems[idx0], rk0)
    assert np.array_equal(to_np(vec.new_problem_idx), idx0)
    assert np.array_equal(M.env.u32_to_numpy(vec.reset_keys), rk0)
    assert np.array_equal(to_np(obs_c), np.stack([obs_r[a] for a in ref.agents], 1))
    assert_state_equal(vec.sat_state(), st_r, "initial ")
    rng_r = key
    arng = np.random.default_rng(1)
    saw_reset = False
    for t in range(7):
        acts = _random_actions(arng, ref, B)
        ks = orollout.rollout_keys(rng_r, B, P)
        rng_r = ks["rng"]
        fo, st_r, rew_r, done_r, info_r = orollout.env_step_with_autoreset(
            ref, st_r, acts, problems, ks["new_problem_indices"], ks["reset_keys"])
        out = vec.step(torch.from_numpy(acts).cuda())
        assert np.array_equal(M.env.u32_to_numpy(vec.keys.chain),
                              np.concatenate([ks["rng"], ks["act_key"], ks["step_key"], ks["prob_key"], ks["reset_key"]]))
        if not fused:
            assert np.array_equal(to_np(vec.new_problem_idx), ks["new_problem_indices"])
            assert np.array_equal(M.env.u32_to_numpy(vec.reset_keys), ks["reset_keys"])
        assert np.array_equal(to_np(out["obs"]), fo), f"step {t} obs"
        assert np.array_equal(to_np(out["reward"]), rew_r)
        assert np.array_equal(to_np(out["done"]).astype(bool), np.repeat(done_r[:, None], ref.num_agents + 1, 1))
        assert np.array_equal(to_np(out["solved"]).astype(bool), info_r["solved"])
        assert np.array_equal(to_np(out["num_unsatisfied"]), info_r["num_unsatisfied"])
        assert np.array_equal(to_np(out["episode_step"]), info_r["episode_step"])
        assert_state_equal(vec.sat_state(), st_r, f"step {t} ")
        saw_reset |= bool(done_r.any())
    assert saw_reset


def test_env_key_derivation_is_shard_invariant():
    M = _msat()
    Bg, P = 1001, 37
    prob_key, reset_key = otf.prng_key(5), otf.prng_key(6)
    exp_idx = otf.randint(prob_key, Bg, 0, P)
    exp_keys = otf.split(reset_key, Bg)
    pk = M.env.as_u32_tensor(prob_key, torch.device("cuda"))
    rk = M.env.as_u32_tensor(reset_key, torch.device("cuda"))
    for world in (1, 2, 3, 8):
        got_idx, got_keys = [], []
        for rank in range(world):
            off, cnt = M.shard_range(Bg, world, rank)
            idx = torch.empty(cnt, dtype=torch.int32, device="cuda")
            keys = torch.empty((cnt, 2), dtype=torch.int32, device="cuda")
            M.derive_env_keys(pk, rk, Bg, off, cnt, P, idx, keys)
            got_idx.append(to_np(idx))
            got_keys.append(M.env.u32_to_numpy(keys))
        assert np.array_equal(np.concatenate(got_idx), exp_idx)
        assert np.array_equal(np.concatenate(got_keys), exp_keys)


@pytest.mark.parametrize("T,B,A", [(1, 1, 1), (16, 33, 5), (129, 257, 3), (512, 1024, 1)])
def test_gae_matches_oracle(T, B, A):
    M = _msat()
    rng = np.random.default_rng(T * 31 + B)
    reward = (rng.random((T, B, A)) < 0.05).astype(np.float32)
    reward[..., 1:] = reward[..., :1]
    done = rng.random((T, B)) < 0.03
    value = rng.standard_normal((T, B)).astype(np.float32)
    last_val = rng.standard_normal((B,)).astype(np.float32)
    adv_r, tgt_r = ogae.calculate_gae(reward, done, value, last_val, 0.995, 0.95)
    dev = "cuda"
    adv_c, tgt_c = M.calculate_gae(torch.from_numpy(reward).to(dev), torch.from_numpy(done).to(dev),
                                   torch.from_numpy(value).to(dev), torch.from_numpy(last_val).to(dev), 0.995, 0.95)
    scale = np.abs(adv_r).max() + 1e-30
    assert np.max(np.abs(to_np(adv_c) - adv_r)) <= 1e-5 * scale          # 1e-5 relative (north_star)
    assert np.max(np.abs(to_np(tgt_c) - tgt_r)) <= 1e-5 * (np.abs(tgt_r).max() + 1e-30)
    # dense team-reward layout [T, B] gives the same result as the strided [T, B, A] view
    adv_d, _ = M.calculate_gae(torch.from_numpy(np.ascontiguousarray(reward[..., 0])).to(dev),
                               torch.from_numpy(done).to(dev), torch.from_numpy(value).to(dev),
                               torch.from_numpy(last_val).to(dev), 0.995, 0.95)
    assert torch.equal(adv_d, adv_c)
    norm_r = ogae.normalize_advantages(adv_r)
    norm_c = M.normalize_advantages(adv_c.clone())
    if T * B > 1:
        assert np.max(np.abs(to_np(norm_c) - norm_r)) <= 1e-5 * (np.abs(norm_r).max() + 1e-30) + 1e-6
    # statistics fused into the scan give the same normalisation as the separate pass
    stats = torch.zeros(3, dtype=torch.float64, device=dev)
    adv_f, _ = M.calculate_gae(torch.from_numpy(reward).to(dev), torch.from_numpy(done).to(dev),
                               torch.from_numpy(value).to(dev), torch.from_numpy(last_val).to(dev), 0.995, 0.95,
                               stats=stats)
    assert torch.equal(adv_f, adv_c)
    ref_stats = M.advantage_stats(adv_c)
    assert float(stats[0]) == T * B and torch.allclose(stats, ref_stats, rtol=1e-12, atol=1e-9)
    if T * B > 1:
        assert torch.allclose(M.normalize_advantages(adv_f.clone(), stats=stats), norm_c, rtol=1e-6, atol=1e-6)


def test_full_size_properties_uf100_65536():
    """BASELINE headline size (uf100-430 x 65,536 envs): size-independent properties instead of a
    full oracle run -- value range, own-variable segment == assignment, num_unsatisfied == m - sum(status),
    solved <=> num_unsatisfied == 0, flip involution (same action twice restores the assignment)."""
    M = _msat()
    n, m, B = 100, 430, 65536
    P = 4096
    problems = torch.from_numpy(_formulas("uniform", P, n, m, 3, seed=20261020))
    env = M.SATEnv(n, m, 512, verbose=False)
    bank = env.make_bank(problems)
    g = torch.Generator(device="cuda").manual_seed(0)
    idx = torch.randint(0, P, (B,), generator=g, device="cuda", dtype=torch.int32)
    keys = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, 2), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    obs, st = env.reset_from_bank(bank, idx, keys)
    assert int(obs.min()) >= -1 and int(obs.max()) <= 1
    assign = st.variable_assignments
    own = torch.stack([obs[:, a, 4 * a:4 * a + 4] for a in range(env.num_agents)], 1).reshape(B, n)
    assert torch.equal(own, assign)
    assert torch.equal(st.num_unsatisfied, (m - st.clauses_satisfied_status.sum(1)).to(torch.int32))
    # subsample against the oracle
    sub = torch.arange(0, B, 1021, device="cuda")
    ref = SATEnvOracle(n, m, 512)
    obs_r, st_r = ref.reset(to_np(problems)[to_np(idx[sub])], M.env.u32_to_numpy(keys[sub]))
    assert np.array_equal(to_np(obs[sub]), np.stack([obs_r[a] for a in ref.agents], 1))
    acts = torch.randint(0, 5, (B, env.num_agents), generator=g, device="cuda", dtype=torch.int32)
    _, st1, rew, done, info = env.step_env(None, st, acts)
    assert torch.equal(info["solved"], info["num_unsatisfied"] == 0)
    assert torch.equal(rew["agent_0"], info["solved"].float())
    _, st2, *_ = env.step_env(None, st1, acts)
    assert torch.equal(st2.variable_assignments, assign)
    assert int(st2.step.min()) == 2 and int(st2.step.max()) == 2


def test_host_buffer_rollout_step_matches_oracle():
    """msat_rollout_step_host (pinned host actions in, compact reward/done/info out) against the oracle."""
    M = _msat()
    n, m, B, P, max_steps = 20, 91, 40, 6, 2
    problems = _formulas("uniform", P, n, m, 3, seed=21)
    ref = SATEnvOracle(n, m, max_steps)
    env = M.SATEnv(n, m, max_steps, verbose=False)
    key0 = otf.prng_key(9)
    vec = M.VecSATEnv(env, problems, B, key0, compact_outputs=True)
    vec.reset()
    key, idx0, rk0 = orollout.initial_reset_inputs(key0, B, P)
    _, st_r = ref.reset(problems[idx0], rk0)
    host = vec.alloc_host_io()
    assert host["reward"].shape == (B, 1) and host["done"].shape == (B, 1) and host["actions"].is_pinned()
    arng = np.random.default_rng(2)
    for t in range(5):
        acts = _random_actions(arng, ref, B)
        ks = orollout.rollout_keys(key, B, P)
        key = ks["rng"]
        fo, st_r, rew_r, done_r, info_r = orollout.env_step_with_autoreset(
            ref, st_r, acts, problems, ks["new_problem_indices"], ks["reset_keys"])
        host["actions"].copy_(torch.from_numpy(acts))
        vec.step_host(host)
        rewards, dones, infos = vec.host_views(host)
        for i, a in enumerate(env.agents):
            assert np.array_equal(rewards[a].numpy(), rew_r[:, i])
            assert np.array_equal(dones[a].numpy(), done_r)
        assert np.array_equal(dones["__all__"].numpy(), done_r)
        assert np.array_equal(infos["solved"].numpy(), info_r["solved"])
        assert np.array_equal(infos["num_unsatisfied"].numpy(), info_r["num_unsatisfied"])
        assert np.array_equal(infos["episode_step"].numpy(), info_r["episode_step"])
        assert np.array_equal(to_np(vec.out["obs"]), fo)
        assert_state_equal(vec.sat_state(), st_r, f"step {t} ")


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rollout_equals_single_device(world):
    """Env sharding: `world` VecSATEnv shards, each deriving its keys from global env
    indices, reproduce the single-device rollout exactly (no data-path collective needed)."""
    M = _msat()
    n, m, B, P, max_steps = 20, 91, 50, 7, 2
    problems = _formulas("uniform", P, n, m, 3, seed=31)
    env = M.SATEnv(n, m, max_steps, verbose=False)
    bank = env.make_bank(problems)
    key0 = otf.prng_key(77)
    full = M.VecSATEnv(env, bank, B, key0)
    shards = [M.VecSATEnv(env, bank, B, key0, world_size=world, rank=r) for r in range(world)]
    obs_full = full.reset()
    obs_sh = torch.cat([s.reset() for s in shards])
    assert torch.equal(obs_full, obs_sh)
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(6):
        acts = torch.randint(0, 5, (B, env.num_agents), generator=g, device="cuda", dtype=torch.int32)
        o = full.step(acts)
        outs = [s.step(acts[s.env_offset:s.env_offset + s.num_envs].contiguous()) for s in shards]
        for k_ in ("obs", "reward", "done", "solved", "num_unsatisfied", "episode_step"):
            assert torch.equal(o[k_], torch.cat([x[k_] for x in outs])), (t, k_)
        assert torch.equal(full.state, torch.cat([s.state for s in shards]))
        for s in shards:
            assert torch.equal(s.keys.chain, full.keys.chain)


@pytest.mark.parametrize("B", [9001, 33001], ids=["single-slice", "four-slices-two-streams"])
def test_host_step_sliced_pipeline_equals_device_step(B):
    """msat_rollout_step_host (coalesced result copy; batches >= 32768 envs sliced over two internal streams)
    must produce exactly the outputs, state and rng chain of the single-launch device step (odd batch sizes:
    ragged last slice)."""
    M = _msat()
    n, m, P, max_steps = 20, 91, 16, 3
    problems = _formulas("uniform", P, n, m, 3, seed=41)
    env = M.SATEnv(n, m, max_steps, verbose=False)
    bank = env.make_bank(problems)
    key0 = otf.prng_key(5)
    dev_vec = M.VecSATEnv(env, bank, B, key0, compact_outputs=True)
    host_vec = M.VecSATEnv(env, bank, B, key0, compact_outputs=True)
    assert torch.equal(dev_vec.reset(), host_vec.reset())
    host = host_vec.alloc_host_io()
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(7):
        acts = torch.randint(0, 5, (B, env.num_agents), generator=g, device="cuda", dtype=torch.int32)
        out = dev_vec.step(acts)
        host["actions"].copy_(acts.cpu())
        host_vec.step_host(host)
        torch.cuda.synchronize()
        for k_ in ("reward", "done", "solved", "num_unsatisfied", "episode_step"):
            assert torch.equal(host[k_], out[k_].cpu()), (t, k_)
            assert torch.equal(host_vec.out[k_], out[k_]), (t, k_)
        assert torch.equal(host_vec.out["obs"], out["obs"])
        assert torch.equal(host_vec.state, dev_vec.state)
        assert torch.equal(host_vec.keys.chain, dev_vec.keys.chain)
