"""GPU parity tests of the round-2 entry points (through the C ABI): K fused steps per launch, the
asynchronous host pipeline, the shaped reward, the cp.async-pipelined GAE scan, the vectorised
normalisation, and oracle comparisons at the BASELINE sizes (sub-sampled)."""
import numpy as np
import pytest
import torch

from oracle import gae as ogae
from oracle import rollout as orollout
from oracle import threefry as otf
from oracle.sat_env import SATEnvOracle
from tests.util import to_np

pytestmark = pytest.mark.gpu


def _msat():
    import marl_sat_b200 as M
    return M


def _formulas(kind, P, n, m, k, seed):
    from marl_sat_b200.synth import mixed_ksat, uniform_ksat
    return uniform_ksat(P, n, m, k, seed) if kind == "uniform" else mixed_ksat(P, n, m, 3, k, seed)


def _actions(rng, env, shape_prefix):
    if env.action_mode == 0:
        return rng.integers(0, env.max_vars_per_agent + 1, size=shape_prefix + (env.num_agents,)).astype(np.int32)
    return rng.integers(0, 2, size=shape_prefix + (env.num_agents, env.max_vars_per_agent)).astype(np.int32)


# ---------------------------------------------------------------------------------------------------------
# msat_rollout_steps: K steps in one launch == K launches of msat_rollout_step
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m,k,vpa,kind,mode,gs,B,K,max_steps", [
    (20, 91, 3, None, "uniform", 0, 0, 77, 9, 3),          # one-warp groups, several resets inside the launch
    (50, 218, 3, None, "uniform", 0, 0, 33, 16, 5),
    (50, 218, 3, None, "uniform", 1, 64, 21, 7, 2),        # 64-thread groups (named barriers), multi-flip
    (100, 430, 3, None, "uniform", 0, 0, 19, 5, 2),        # 256-thread groups
    (100, 430, 7, 7, "mixed", 0, 128, 11, 6, 3),           # padding quirk shape, 128-thread groups
    (12, 20, 3, None, "uniform", 0, 0, 64, 64, 6),         # K at its maximum, frequent solves
    (20, 91, 3, None, "uniform", 0, 16, 77, 9, 3),         # half-warp groups: two envs per warp, resets diverge
    (12, 20, 3, None, "uniform", 1, 16, 35, 33, 4),
    (20, 91, 3, None, "uniform", 0, 32, 77, 9, 3),         # the same shape with one-warp groups
    (35, 149, 3, 7, "uniform", 1, 16, 41, 12, 4),
])
@pytest.mark.parametrize("every", [True, False])
def test_multi_step_launch_equals_single_steps(n, m, k, vpa, kind, mode, gs, B, K, max_steps, every):
    M = _msat()
    P = 9
    problems = _formulas(kind, P, n, m, k, seed=n + K)
    env = M.SATEnv(n, m, max_steps, vars_per_agent=vpa, action_mode=mode, verbose=False, group_threads=gs)
    bank = env.make_bank(problems)
    key0 = otf.prng_key(3 + K)
    one = M.VecSATEnv(env, bank, B, key0)
    many = M.VecSATEnv(env, bank, B, key0)
    assert torch.equal(one.reset(), many.reset())
    acts = torch.from_numpy(_actions(np.random.default_rng(K), env, (K, B))).cuda()
    out = many.alloc_multi_step_outputs(K, emit_every_step=every)
    many.steps(acts, out)
    resets = 0
    for j in range(K):
        o = one.step(acts[j])
        for name in ("reward", "done", "solved", "num_unsatisfied", "episode_step"):
            assert torch.equal(out[name][j], o[name]), (j, name)
        if every:
            assert torch.equal(out["obs"][j], o["obs"]), j
        resets += int(o["done"][:, -1].sum())
    if not every:
        assert torch.equal(out["obs"], one.out["obs"])
    assert torch.equal(many.state, one.state)
    assert torch.equal(many.keys.chain, one.keys.chain)
    assert resets > 0


@pytest.mark.parametrize("every", [True, False])
def test_multi_step_gnn_mode_and_shaped_reward(every):
    M = _msat()
    n, m, B, K, P = 35, 149, 45, 12, 5
    problems = _formulas("uniform", P, n, m, 3, seed=8)
    env = M.SATEnv(n, m, 4, vars_per_agent=7, verbose=False, reward_mode="shaped", r_clause=0.05, r_sat=20.0, gamma=0.995)
    bank = env.make_bank(problems)
    key0 = otf.prng_key(12)
    one = M.VecSATEnv(env, bank, B, key0, emit_obs=False, gnn_outputs=True)
    many = M.VecSATEnv(env, bank, B, key0, emit_obs=False, gnn_outputs=True)
    one.reset(); many.reset()
    acts = torch.from_numpy(_actions(np.random.default_rng(1), env, (K, B))).cuda()
    out = many.alloc_multi_step_outputs(K, emit_every_step=every)
    many.steps(acts, out)
    for j in range(K):
        o = one.step(acts[j])
        for name in ("reward", "done", "solved", "num_unsatisfied", "episode_step"):
            assert torch.equal(out[name][j], o[name]), (j, name)
        if every:
            assert torch.equal(out["gnn_assignment"][j], o["gnn_assignment"])
            assert torch.equal(out["gnn_clause_features"][j], o["gnn_clause_features"])
    if not every:
        assert torch.equal(out["gnn_assignment"], one.out["gnn_assignment"])
        assert torch.equal(out["gnn_clause_features"], one.out["gnn_clause_features"])
    assert torch.equal(many.state, one.state)
    assert out["reward"].unique().numel() > 4          # shaped, not 0/1


# ---------------------------------------------------------------------------------------------------------
# shaped reward (env:201-223) against the oracle, incl. newly_satisfied
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
def test_shaped_reward_matches_oracle(mode):
    M = _msat()
    n, m, B = 23, 60, 31
    kw = dict(r_clause=0.02, r_sat=1.0, gamma=0.99, reward_mode="shaped")
    problems = _formulas("uniform", B, n, m, 3, seed=4)
    keys = np.random.default_rng(5).integers(0, 2 ** 32, size=(B, 2), dtype=np.uint64).astype(np.uint32)
    ref = SATEnvOracle(n, m, 6, action_mode=mode, **kw)
    env = M.SATEnv(n, m, 6, action_mode=mode, verbose=False, **kw)
    _, st_r = ref.reset(problems, keys)
    _, st_c = env.reset(problems, keys)
    rng = np.random.default_rng(6)
    for t in range(8):
        acts = _actions(rng, ref, (B,))
        _, st_r, rew_r, _, info_r = ref.step_env(None, st_r, acts)
        _, st_c, rew_c, _, info_c = env.step_env(None, st_c, acts)
        for a in env.agents:
            assert np.array_equal(to_np(rew_c[a]), rew_r[a]), (t, a)
        assert np.array_equal(to_np(info_c["newly_satisfied"]), info_r["newly_satisfied"])


# ---------------------------------------------------------------------------------------------------------
# asynchronous host pipeline == device step
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [4097, 20000])
def test_async_host_pipeline_equals_device_step(B):
    M = _msat()
    n, m, P, max_steps, T = 20, 91, 16, 3, 11
    problems = _formulas("uniform", P, n, m, 3, seed=41)
    env = M.SATEnv(n, m, max_steps, verbose=False)
    bank = env.make_bank(problems)
    key0 = otf.prng_key(5)
    dev_vec = M.VecSATEnv(env, bank, B, key0, compact_outputs=True)
    host_vec = M.VecSATEnv(env, bank, B, key0, compact_outputs=True)
    assert torch.equal(dev_vec.reset(), host_vec.reset())
    slots = host_vec.alloc_async_io(depth=2)
    acts_np = _actions(np.random.default_rng(3), env, (T, B))
    table = torch.from_numpy(acts_np).pin_memory()
    expected = []
    for t in range(T):
        o = dev_vec.step(torch.from_numpy(acts_np[t]).cuda())
        expected.append({k_: o[k_].cpu().clone() for k_ in ("reward", "done", "solved", "num_unsatisfied", "episode_step")})
    got = [None] * T

    def harvest(t):
        host_vec.host_wait(t % 2)
        got[t] = {k_: slots[t % 2]["host"][k_].clone() for k_ in expected[0]}
    for t in range(T):
        if t >= 2:
            harvest(t - 2)            # the host reads step t-2 while steps t-1 / t are in flight
        host_vec.step_host_async(t % 2, slots[t % 2], actions_host=table[t])
    harvest(T - 2)
    harvest(T - 1)
    for t in range(T):
        for k_ in expected[t]:
            assert torch.equal(got[t][k_], expected[t][k_]), (t, k_)
    torch.cuda.synchronize()
    assert torch.equal(host_vec.state, dev_vec.state)
    assert torch.equal(host_vec.keys.chain, dev_vec.keys.chain)
    assert torch.equal(host_vec.out["obs"], dev_vec.out["obs"])
    host_vec.close()


# ---------------------------------------------------------------------------------------------------------
# GAE: cp.async-pipelined plain scan (large batches) and the vectorised normalisation
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,B,strided", [(37, 40000, False), (64, 37904, False), (9, 65536, True), (33, 40008, False),
                                         (512, 38000, False)])
def test_gae_pipelined_scan_matches_oracle(T, B, strided):
    """B >= 37,888 columns fill the GPU -> S = 1; with 16-byte copyable rows the pipelined kernel runs
    (B % 16 == 0, dense team reward), otherwise the register-chunked one.  Both against the oracle, and
    against each other bit for bit (same operation order)."""
    M = _msat()
    from marl_sat_b200 import _lib
    rng = np.random.default_rng(T + B)
    reward = (rng.random((T, B)) < 0.05).astype(np.float32)
    done = rng.random((T, B)) < 0.03
    value = rng.standard_normal((T, B)).astype(np.float32)
    last_val = rng.standard_normal((B,)).astype(np.float32)
    adv_r, tgt_r = ogae.calculate_gae(reward, done, value, last_val, 0.995, 0.95)
    r_t = torch.from_numpy(reward).cuda()
    if strided:
        r_t = r_t[:, :, None].expand(T, B, 3).contiguous()
    args = (r_t, torch.from_numpy(done).cuda(), torch.from_numpy(value).cuda(), torch.from_numpy(last_val).cuda())
    stats = torch.zeros(3, dtype=torch.float64, device="cuda")
    adv_c, tgt_c = M.calculate_gae(*args, 0.995, 0.95, stats=stats)
    assert np.array_equal(to_np(adv_c), adv_r) or np.max(np.abs(to_np(adv_c) - adv_r)) <= 1e-5 * np.abs(adv_r).max()
    assert np.max(np.abs(to_np(tgt_c) - tgt_r)) <= 1e-5 * np.abs(tgt_r).max()
    lib = _lib.load()
    assert lib.msat_tune(b"gae_plain", 1) == 0
    try:
        stats_p = torch.zeros(3, dtype=torch.float64, device="cuda")
        adv_p, tgt_p = M.calculate_gae(*args, 0.995, 0.95, stats=stats_p)
    finally:
        lib.msat_tune(b"gae_plain", 0)
    assert torch.equal(adv_p, adv_c) and torch.equal(tgt_p, tgt_c)
    for width in (4, 2, 1):            # every column width of the pipelined kernel (normally chosen by batch size)
        lib.msat_tune(b"gae_variant", width)
        try:
            adv_w, tgt_w = M.calculate_gae(*args, 0.995, 0.95)
        finally:
            lib.msat_tune(b"gae_variant", 0)
        assert torch.equal(adv_w, adv_c) and torch.equal(tgt_w, tgt_c), width
    assert float(stats[0]) == T * B and torch.allclose(stats, stats_p, rtol=1e-12, atol=1e-7)
    norm_r = ogae.normalize_advantages(adv_r)
    norm_c = M.normalize_advantages(adv_c.clone(), stats=stats)
    assert np.max(np.abs(to_np(norm_c) - norm_r)) <= 1e-5 * np.abs(norm_r).max() + 1e-6


@pytest.mark.parametrize("count,offset", [(1, 0), (3, 1), (4, 0), (5, 3), (1000, 1), (4099, 2), (1 << 20, 0), ((1 << 20) + 7, 3)])
def test_normalize_any_alignment(count, offset):
    M = _msat()
    from marl_sat_b200 import _lib
    from marl_sat_b200.env import _ptr, _stream_ptr
    base = torch.randn(count + 8, device="cuda")
    guard = base.clone()
    view = base[offset:offset + count]
    stats = M.advantage_stats(view.contiguous())
    _lib.check(_lib.load().msat_adv_normalize(_ptr(view), count, _ptr(stats), _stream_ptr(view.device)), "norm")
    x = guard[offset:offset + count].double()
    mean, std = x.mean(), x.var(unbiased=False).sqrt() if count > 1 else torch.tensor(0.0, dtype=torch.float64)
    exp = ((guard[offset:offset + count] - mean.float()) / (std.float() + 1e-8))
    assert torch.allclose(view, exp, rtol=1e-5, atol=1e-5)
    assert torch.equal(base[:offset], guard[:offset]) and torch.equal(base[offset + count:], guard[offset + count:])


def test_normalize_advantages_non_contiguous_is_in_place():
    """a strided view must be normalised in place (or rejected), never silently copied."""
    M = _msat()
    buf = torch.randn(16, 8, 3, device="cuda")
    view = buf[:, :, 0]
    ref = (view - view.mean()) / (view.std(unbiased=False) + 1e-8)
    out = M.normalize_advantages(view)
    assert torch.allclose(view, ref, rtol=1e-5, atol=1e-5)
    assert out.data_ptr() == view.data_ptr()


# ---------------------------------------------------------------------------------------------------------
# BASELINE sizes: step against the oracle on a sub-sample (the C port checks whole sub-batches quickly)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,m,k,vpa,kind,B,stride", [
    ("c3_uf100_65536", 100, 430, 3, None, "uniform", 65536, 509),
    ("c4_uf250_16384", 250, 1065, 3, None, "uniform", 16384, 257),
    ("c5_mixedk_32768", 100, 430, 7, 7, "mixed", 32768, 251),
    ("c2_uf50_4096", 50, 218, 3, None, "uniform", 4096, 31),
])
def test_baseline_sizes_rollout_subsample_matches_oracle(name, n, m, k, vpa, kind, B, stride):
    """Three rollout steps with auto-reset (max_steps = 2 -> every env resets at step 2) at the BASELINE batch
    sizes; the envs ``0, stride, 2*stride, ...`` are replayed by the oracle with the same global keys."""
    M = _msat()
    P, max_steps, T = 64, 2, 3
    problems = _formulas(kind, P, n, m, k, seed=B % 97)
    env = M.SATEnv(n, m, max_steps, vars_per_agent=vpa, verbose=False)
    ref = SATEnvOracle(n, m, max_steps, vars_per_agent=vpa)
    key0 = otf.prng_key(B)
    vec = M.VecSATEnv(env, problems, B, key0)
    obs = vec.reset()
    sub = np.arange(0, B, stride)
    key, idx0, rk0 = orollout.initial_reset_inputs(key0, B, P)
    obs_r, st_r = ref.reset(problems[idx0[sub]], rk0[sub])
    assert np.array_equal(to_np(obs[sub]), np.stack([obs_r[a] for a in ref.agents], 1))
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(T):
        acts = torch.randint(0, env.max_vars_per_agent + 1, (B, env.num_agents), generator=g, device="cuda",
                             dtype=torch.int32)
        o = vec.step(acts)
        ks = orollout.rollout_keys(key, B, P)
        key = ks["rng"]
        fo, st_r, rew_r, done_r, info_r = orollout.env_step_with_autoreset(
            ref, st_r, to_np(acts)[sub], problems, ks["new_problem_indices"][sub], ks["reset_keys"][sub])
        assert np.array_equal(to_np(o["obs"][sub]), fo), t
        assert np.array_equal(to_np(o["reward"][sub]), rew_r), t
        assert np.array_equal(to_np(o["done"][sub][:, -1]).astype(bool), done_r), t
        assert np.array_equal(to_np(o["num_unsatisfied"][sub]), info_r["num_unsatisfied"]), t
        assert np.array_equal(to_np(o["episode_step"][sub]), info_r["episode_step"]), t
    assert int(o["episode_step"].max()) == 1 and int(o["done"].sum()) == 0      # everyone restarted after step 2
    st = vec.sat_state()
    assert np.array_equal(to_np(st.variable_assignments[sub]), st_r.variable_assignments)


@pytest.mark.parametrize("obs_dtype", [torch.int32, torch.int8])
def test_sliced_host_step_at_headline_shape(obs_dtype):
    """msat_rollout_step_host's four-slice two-stream path on uf100-430 (the headline shape, 256-thread groups;
    with int8 observations the slices' observation offsets are in bytes)."""
    M = _msat()
    n, m, B, P = 100, 430, 33001, 32
    problems = _formulas("uniform", P, n, m, 3, seed=2)
    env = M.SATEnv(n, m, 3, verbose=False, obs_dtype=obs_dtype)
    bank = env.make_bank(problems)
    key0 = otf.prng_key(8)
    dev_vec = M.VecSATEnv(env, bank, B, key0, compact_outputs=True)
    host_vec = M.VecSATEnv(env, bank, B, key0, compact_outputs=True)
    dev_vec.reset(); host_vec.reset()
    host = host_vec.alloc_host_io()
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(4):
        acts = torch.randint(0, 5, (B, env.num_agents), generator=g, device="cuda", dtype=torch.int32)
        out = dev_vec.step(acts)
        host["actions"].copy_(acts.cpu())
        host_vec.step_host(host)
        for k_ in ("reward", "done", "solved", "num_unsatisfied", "episode_step"):
            assert torch.equal(host[k_], out[k_].cpu()), (t, k_)
        assert torch.equal(host_vec.out["obs"], out["obs"])
    assert torch.equal(host_vec.state, dev_vec.state)


def test_problem_index_validation():
    M = _msat()
    env = M.SATEnv(20, 91, 5, verbose=False)
    bank = env.make_bank(_formulas("uniform", 4, 20, 91, 3, seed=1))
    keys = np.zeros((3, 2), np.uint32)
    with pytest.raises(IndexError):
        env.reset_from_bank(bank, torch.tensor([0, 4, 1], dtype=torch.int32), keys, validate_indices=True)
    with pytest.raises(IndexError):
        env.reset_from_bank(bank, torch.tensor([0, -1, 1], dtype=torch.int32), keys, validate_indices=True)
    env.reset_from_bank(bank, torch.tensor([0, 3, 1], dtype=torch.int32), keys, validate_indices=True)


# ---------------------------------------------------------------------------------------------------------
# MSAT_OBS_INT8: the same observations in one byte per element
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m,k,vpa,kind,gs,B", [
    (20, 91, 3, None, "uniform", 0, 77),        # half-warp groups, 655-element rows (odd: every 16-byte phase)
    (20, 91, 3, None, "uniform", 32, 5),
    (7, 12, 3, None, "uniform", 0, 33),         # rows shorter than one 16-byte chunk pair
    (3, 2, 2, 1, "uniform", 0, 19),             # an env's whole range is shorter than one chunk
    (50, 218, 3, None, "uniform", 0, 41),
    (50, 218, 3, None, "uniform", 64, 9),
    (100, 430, 7, 7, "mixed", 0, 13),           # 128-thread groups
    (100, 430, 3, None, "uniform", 0, 11),      # 256-thread groups
])
def test_int8_observations_equal_int32(n, m, k, vpa, kind, gs, B):
    M = _msat()
    P, K = 6, 5
    problems = _formulas(kind, P, n, m, k, seed=n + B)
    key0 = otf.prng_key(n)
    envs = [M.SATEnv(n, m, 3, vars_per_agent=vpa, verbose=False, group_threads=gs, obs_dtype=dt)
            for dt in (torch.int32, torch.int8)]
    vecs = [M.VecSATEnv(e, problems, B, key0) for e in envs]
    o32, o8 = (v.reset() for v in vecs)
    assert o8.dtype == torch.int8 and o32.dtype == torch.int32 and torch.equal(o8.to(torch.int32), o32)
    rng = np.random.default_rng(B)
    for t in range(4):                           # max_steps = 3: auto-resets inside
        acts = torch.from_numpy(_actions(rng, envs[0], (B,))).cuda()
        a, b = (v.step(acts) for v in vecs)
        assert torch.equal(b["obs"].to(torch.int32), a["obs"]), t
        assert torch.equal(a["reward"], b["reward"]) and torch.equal(a["done"], b["done"])
    table = torch.from_numpy(_actions(rng, envs[0], (K, B))).cuda()
    outs = [v.alloc_multi_step_outputs(K, emit_every_step=True) for v in vecs]
    for v, o in zip(vecs, outs):
        v.steps(table, o)
    assert outs[1]["obs"].dtype == torch.int8 and torch.equal(outs[1]["obs"].to(torch.int32), outs[0]["obs"])
    assert torch.equal(vecs[0].state, vecs[1].state)
    g32, g8 = (e.get_obs_array(v.sat_state()) for e, v in zip(envs, vecs))
    assert g8.dtype == torch.int8 and torch.equal(g8.to(torch.int32), g32)
