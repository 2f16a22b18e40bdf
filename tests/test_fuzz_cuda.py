"""Seeded differential fuzzing of the CUDA path against the oracle over random shapes: odd sizes that
are not multiples of 32, observation rows shorter than one mask word, more than 32 agents (multi-word
agent sets), clauses with repeated variables and all-padding clauses, single-literal clauses, tiny and
empty batches, max_steps = 1, both action modes, every thread-group size."""
import numpy as np
import pytest
import torch

from oracle import rollout as orollout
from oracle import threefry as otf
from oracle.sat_env import SATEnvOracle
from tests.util import assert_obs_equal, assert_state_equal, to_np

pytestmark = pytest.mark.gpu


def _random_formulas(rng, B, n, m, k):
    """Unstructured literals in [-n, n]: repeated variables, x and -x together, zeros anywhere."""
    cl = rng.integers(-n, n + 1, size=(B, m, k)).astype(np.int32)
    cl[rng.random((B, m)) < 0.05] = 0                      # some all-padding clauses
    return cl


@pytest.mark.parametrize("clause_update", ["full", "incremental"])
@pytest.mark.parametrize("seed", range(24))
def test_random_shapes(seed, clause_update):
    import marl_sat_b200 as M
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 90))
    m = int(rng.integers(1, 200))
    k = int(rng.integers(1, 6))
    vpa = [None, 1, 2, 3, 5, 7][int(rng.integers(0, 6))]
    if vpa is not None and vpa > n:
        vpa = n
    mode = int(rng.integers(0, 2))
    B = int(rng.integers(1, 40))
    gs = [0, 32, 64, 128, 256, 16][seed % 6]
    max_steps = int(rng.integers(1, 4))
    cl = _random_formulas(rng, B, n, m, k)
    keys = rng.integers(0, 2 ** 32, size=(B, 2), dtype=np.uint64).astype(np.uint32)
    ref = SATEnvOracle(n, m, max_steps, vars_per_agent=vpa, action_mode=mode)
    env = M.SATEnv(n, m, max_steps, vars_per_agent=vpa, action_mode=mode, verbose=False, group_threads=gs,
                   clause_update=clause_update)
    obs_r, st_r = ref.reset(cl, keys)
    obs_c, st_c = env.reset(cl, keys)
    what = f"[n={n} m={m} k={k} vpa={vpa} mode={mode} B={B} gs={gs}] "
    assert_obs_equal(obs_c, obs_r, env.agents, what + "reset ")
    assert_state_equal(st_c, st_r, what + "reset ")
    A, V = ref.num_agents, ref.max_vars_per_agent
    for t in range(3):
        acts = (rng.integers(-1, V + 2, size=(B, A)) if mode == 0 else rng.integers(0, 2, size=(B, A, V))).astype(np.int32)
        obs_r, st_r, rew_r, done_r, info_r = ref.step_env(None, st_r, acts)
        obs_c, st_c, rew_c, done_c, info_c = env.step_env(None, st_c, acts)
        assert_obs_equal(obs_c, obs_r, env.agents, what + f"step {t} ")
        assert_state_equal(st_c, st_r, what + f"step {t} ")
        assert np.array_equal(to_np(rew_c[env.agents[-1]]), rew_r[ref.agents[-1]])
        assert np.array_equal(to_np(done_c["__all__"]), done_r["__all__"])
        assert np.array_equal(to_np(info_c["num_unsatisfied"]), info_r["num_unsatisfied"])


@pytest.mark.parametrize("clause_update", ["full", "incremental"])
@pytest.mark.parametrize("seed", range(6))
def test_random_rollouts_with_autoreset(seed, clause_update):
    import marl_sat_b200 as M
    rng = np.random.default_rng(77 + seed)
    n, m = int(rng.integers(4, 60)), int(rng.integers(5, 120))
    B, P = int(rng.integers(1, 50)), int(rng.integers(1, 9))
    vpa = [None, 3, 6][seed % 3]
    max_steps = int(rng.integers(1, 4))
    problems = _random_formulas(rng, P, n, m, 3)
    ref = SATEnvOracle(n, m, max_steps, vars_per_agent=vpa)
    env = M.SATEnv(n, m, max_steps, vars_per_agent=vpa, verbose=False, clause_update=clause_update)
    key0 = otf.prng_key(seed)
    vec = M.VecSATEnv(env, problems, B, key0)
    obs_c = vec.reset()
    key, idx0, rk0 = orollout.initial_reset_inputs(key0, B, P)
    obs_r, st_r = ref.reset(problems[idx0], rk0)
    assert np.array_equal(to_np(obs_c), np.stack([obs_r[a] for a in ref.agents], 1))
    for t in range(6):
        acts = rng.integers(0, ref.max_vars_per_agent + 1, size=(B, ref.num_agents)).astype(np.int32)
        ks = orollout.rollout_keys(key, B, P)
        key = ks["rng"]
        fo, st_r, rew_r, done_r, info_r = orollout.env_step_with_autoreset(
            ref, st_r, acts, problems, ks["new_problem_indices"], ks["reset_keys"])
        out = vec.step(torch.from_numpy(acts).cuda())
        assert np.array_equal(to_np(out["obs"]), fo), f"step {t}"
        assert np.array_equal(to_np(out["done"][:, -1]).astype(bool), done_r)
        assert_state_equal(vec.sat_state(), st_r, f"step {t} ")


def test_empty_batch_is_a_no_op():
    import marl_sat_b200 as M
    env = M.SATEnv(20, 91, 5, verbose=False)
    from marl_sat_b200.synth import uniform_ksat
    bank = env.make_bank(uniform_ksat(3, 20, 91, 3, 0))
    idx = torch.empty((0,), dtype=torch.int32, device="cuda")
    keys = torch.empty((0, 2), dtype=torch.int32, device="cuda")
    obs, st = env.reset_from_bank(bank, idx, keys)
    assert obs.shape == (0, 5, 131) and st.num_envs == 0
    adv, tgt = M.calculate_gae(torch.empty((0, 4), device="cuda"), torch.empty((0, 4), dtype=torch.uint8, device="cuda"),
                               torch.empty((0, 4), device="cuda"), torch.zeros(4, device="cuda"), 0.9, 0.9)
    assert adv.shape == (0, 4)


@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("gs", [0, 16, 32, 256])
def test_timeout_boundary_stress(gs):
    """Many envs, tiny max_steps: every env crosses the time-out boundary again and again, so every CTA
    repeatedly takes the 'one step before time-out' and the reset path.  Regression test for a
    shared-memory race on the step counter (a warp that re-read `step` after thread 0 had advanced it
    disagreed about `done` and dead-locked the group barrier): all envs are checked with invariants, a
    subset against the oracle."""
    import marl_sat_b200 as M
    from marl_sat_b200.synth import uniform_ksat
    n, m, B, P, max_steps = 20, 91, 16384, 32, 3
    problems = uniform_ksat(P, n, m, 3, seed=3)
    ref = SATEnvOracle(n, m, max_steps)
    env = M.SATEnv(n, m, max_steps, verbose=False, group_threads=gs)
    key0 = otf.prng_key(123)
    vec = M.VecSATEnv(env, problems, B, key0)
    vec.reset()
    key, idx0, rk0 = orollout.initial_reset_inputs(key0, B, P)
    sub = np.arange(0, B, 67)
    _, st_r = ref.reset(problems[idx0[sub]], rk0[sub])
    g = torch.Generator(device="cuda").manual_seed(5)
    for t in range(14):
        acts = torch.randint(0, 5, (B, env.num_agents), generator=g, device="cuda", dtype=torch.int32)
        out = vec.step(acts)
        ks = orollout.rollout_keys(key, B, P)
        key = ks["rng"]
        fo, st_r, rew_r, done_r, info_r = orollout.env_step_with_autoreset(
            ref, st_r, to_np(acts)[sub], problems, ks["new_problem_indices"][sub], ks["reset_keys"][sub])
        es, done, solved = out["episode_step"], out["done"][:, -1].bool(), out["solved"].bool()
        assert int(es.min()) >= 1 and int(es.max()) <= max_steps
        assert torch.equal(done, solved | (es == max_steps))
        assert np.array_equal(to_np(out["obs"])[sub], fo), f"step {t}"
        assert np.array_equal(to_np(done)[sub], done_r)
        assert np.array_equal(to_np(es)[sub], info_r["episode_step"])
    torch.cuda.synchronize()


@pytest.mark.parametrize("variant", ["incremental", "full_k3"])
@pytest.mark.parametrize("seed", range(16))
def test_incremental_clause_update_without_observations(seed, variant):
    """The launches that write no observations have their own clause evaluators: the incremental (CSR
    occurrence-list) update (``variant == "incremental"``) and, for k = 3 with the full recompute, the paired
    evaluator (two adjacent clauses per lane, no status words; ``variant == "full_k3"``: odd and even m, more than
    one 256-clause staging pass, every 16-byte phase of the feature rows).  Rollouts through ``emit_obs=False`` /
    ``gnn_outputs=True`` -- single steps and K fused steps, interleaved with observation-writing launches on the
    same state -- against the oracle, on unstructured formulas (repeated variables, x and -x in one clause, 0
    padding anywhere), both action modes, every group size."""
    import marl_sat_b200 as M
    from oracle import features as ofeat
    rng = np.random.default_rng(4000 + seed)
    n, m, k = int(rng.integers(2, 70)), int(rng.integers(1, 300)), int(rng.integers(1, 8))
    if variant == "full_k3":
        m, k = int(rng.integers(1, 700)), 3
    vpa = [None, 1, 3, 5, 7][seed % 5]
    if vpa is not None and vpa > n:
        vpa = n
    mode = seed % 2
    B, P = int(rng.integers(1, 40)), int(rng.integers(1, 7))
    gs = [0, 32, 64, 128, 256, 16][(seed // 2) % 6]
    max_steps = int(rng.integers(1, 5))
    problems = _random_formulas(rng, P, n, m, k)
    ref = SATEnvOracle(n, m, max_steps, vars_per_agent=vpa, action_mode=mode)
    env = M.SATEnv(n, m, max_steps, vars_per_agent=vpa, action_mode=mode, verbose=False, group_threads=gs,
                   clause_update="incremental" if variant == "incremental" else "full")
    key0 = otf.prng_key(seed)
    gnn = seed % 3 != 0
    vec = M.VecSATEnv(env, problems, B, key0, emit_obs=False, gnn_outputs=gnn)
    vec.reset()
    key, idx0, rk0 = orollout.initial_reset_inputs(key0, B, P)
    _, st_r = ref.reset(problems[idx0], rk0)
    A, V = ref.num_agents, ref.max_vars_per_agent
    what = f"[n={n} m={m} k={k} vpa={vpa} mode={mode} B={B} gs={gs} gnn={gnn}] "

    def draw(shape_prefix):
        return (rng.integers(-1, V + 2, size=shape_prefix + (A,)) if mode == 0
                else rng.integers(0, 2, size=shape_prefix + (A, V))).astype(np.int32)

    def oracle_step(acts):
        nonlocal key, st_r
        ks = orollout.rollout_keys(key, B, P)
        key = ks["rng"]
        fo, st_r, rew, done, info = orollout.env_step_with_autoreset(ref, st_r, acts, problems,
                                                                     ks["new_problem_indices"], ks["reset_keys"])
        return fo, done, info

    for t in range(5):
        acts = draw((B,))
        fo, done_r, info_r = oracle_step(acts)
        out = vec.step(torch.from_numpy(acts).cuda())
        assert np.array_equal(to_np(out["done"][:, -1]).astype(bool), done_r), what + f"step {t}"
        assert np.array_equal(to_np(out["num_unsatisfied"]), info_r["num_unsatisfied"]), what + f"step {t}"
        assert_state_equal(vec.sat_state(), st_r, what + f"step {t} ")
        if gnn:
            gs_r = ofeat.state_to_gnn_input(ref, st_r)
            assert np.array_equal(to_np(out["gnn_assignment"]), gs_r["assignment"])
            assert np.array_equal(to_np(out["gnn_clause_features"]), gs_r["clause_features"])
        # an observation-writing launch in between (full evaluation) must leave the counts consistent
        if t == 2:
            assert np.array_equal(to_np(env.get_obs_array(vec.sat_state())), fo)
            acts = draw((B,))
            fo, done_r, info_r = oracle_step(acts)
            o2 = env.alloc_step_outputs(B, vec.bank.plan.dims)
            vec.keys.advance()              # the un-fused path: chain kernel + key derivation + msat_step with obs
            M.derive_env_keys(vec.keys.prob_key, vec.keys.reset_key, B, 0, B, P, vec.new_problem_idx, vec.reset_keys)
            env.step_into(vec.bank, vec.state, vec.state, torch.from_numpy(acts).cuda(), o2, auto_reset=True,
                          new_problem_idx=vec.new_problem_idx, reset_keys=vec.reset_keys)
            assert np.array_equal(to_np(o2["obs"]), fo), what + "obs-writing step"
            assert_state_equal(vec.sat_state(), st_r, what + "after the obs-writing step ")
    # K fused steps
    K = 6
    table = draw((K, B))
    outs = vec.alloc_multi_step_outputs(K, emit_every_step=True)
    vec.steps(torch.from_numpy(table).cuda(), outs)
    for j in range(K):
        fo, done_r, info_r = oracle_step(table[j])
        assert np.array_equal(to_np(outs["done"][j][:, -1]).astype(bool), done_r), what + f"fused step {j}"
        assert np.array_equal(to_np(outs["num_unsatisfied"][j]), info_r["num_unsatisfied"]), what + f"fused step {j}"
        if gnn:
            gs_r = ofeat.state_to_gnn_input(ref, st_r)
            assert np.array_equal(to_np(outs["gnn_clause_features"][j]), gs_r["clause_features"]), what + f"fused {j}"
    assert_state_equal(vec.sat_state(), st_r, what + "after the fused steps ")
    assert np.array_equal(to_np(env.get_obs_array(vec.sat_state())), fo)
