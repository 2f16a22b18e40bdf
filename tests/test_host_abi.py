"""CPU tests of the host layer: the C ABI library loads and exports every symbol declared in
include/marl_sat_b200.h, plan/dims arithmetic, the drop-in SATEnv's constructor-time attributes,
spaces, sharding arithmetic, and loud failure (no CPU fallback).  No kernel is launched here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import marl_sat_b200 as M
from marl_sat_b200 import _lib
from oracle.sat_env import SATEnvOracle

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "marl_sat_b200.h").read_text()
    declared = set(re.findall(r"\b(msat_[a-z0-9_]+)\s*\(", header))
    assert {"msat_step", "msat_reset", "msat_compile_bank", "msat_gae", "msat_env_keys"} <= declared
    lib = C.CDLL(str(_lib.lib_path()))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert set(_lib.SIGNATURES) <= declared, sorted(set(_lib.SIGNATURES) - declared)
    assert b"sm_100a" in _lib.load().msat_version()


@pytest.mark.parametrize("n,m,vpa", [(20, 91, None), (35, 149, 7), (50, 218, None), (100, 430, None),
                                     (250, 1065, None), (100, 430, 7), (7, 4, 4), (3, 5, None)])
def test_constructor_matches_reference_grouping(n, m, vpa):
    ref = SATEnvOracle(n, m, 10, vars_per_agent=vpa)
    env = M.SATEnv(n, m, 10, vars_per_agent=vpa, verbose=False, device="cpu")
    assert env.agents == ref.agents and env.agent_groups == ref.agent_groups
    assert env.num_agents == ref.num_agents and env.max_vars_per_agent == ref.max_vars_per_agent
    assert np.array_equal(env.agent_vars.numpy(), ref.agent_vars)
    assert np.array_equal(env.action_mask.numpy(), ref.action_mask)
    assert np.array_equal(env.variable_to_agent_idx.numpy(), ref.variable_to_agent_idx)
    assert env.obs_dim == 2 * n + m and env.name == "SATEnv"
    sp = env.action_space("agent_0")
    assert sp.n == env.max_vars_per_agent + 1 and sp.dtype == torch.int32
    ob = env.observation_space("agent_0")
    assert ob.shape == (2 * n + m,) and ob.low == -1 and ob.high == 1 and ob.dtype == torch.float32
    env1 = M.SATEnv(n, m, 10, vars_per_agent=vpa, action_mode=1, verbose=False, device="cpu")
    assert env1.action_space("agent_0").num_categories.tolist() == [2] * env.max_vars_per_agent


def test_constructor_prints_like_the_reference(capsys):
    M.SATEnv(20, 91, 10, device="cpu")
    out = capsys.readouterr().out
    assert "Auto-distribution mode" in out and "Found ideal grouping: 5 agents, each with 4 vars." in out
    M.SATEnv(35, 149, 10, vars_per_agent=7, device="cpu")
    assert "User specified mode: aiming for 7 vars per agent." in capsys.readouterr().out


def test_plan_dims():
    env = M.SATEnv(100, 430, 512, verbose=False, device="cpu")
    d = env._plan_for(3).dims
    assert (d.n, d.m, d.k, d.A, d.V, d.D) == (100, 430, 3, 25, 4, 630)
    assert d.rec_bytes % 128 == 0 and d.rec_bytes >= 430 * 3 * 2 + (25 * 630 + 7) // 8
    assert d.state_words % 4 == 0 and d.state_words >= 4 + 4
    assert d.group_threads in (16, 32, 64, 128, 256) and 0 < d.smem_bytes <= 227 * 1024
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.msat_plan_create(C.byref(h), 0, 1, 3, 1, 0, 1, 0) == _lib.MSAT_EINVAL
    assert lib.msat_plan_create(C.byref(h), 10, 5, 3, 11, 0, 1, 0) == _lib.MSAT_EINVAL       # A > n
    assert lib.msat_plan_create(C.byref(h), 10, 5, 3, 2, 2, 1, 0) == _lib.MSAT_EINVAL        # bad mode
    assert lib.msat_plan_create(C.byref(h), 10, 5, 3, 2, 0, 1, 48) == _lib.MSAT_EINVAL       # bad group size


def test_argument_errors_enqueue_nothing():
    lib = _lib.load()
    env = M.SATEnv(20, 91, 8, verbose=False, device="cpu")
    plan = env._plan_for(3).handle
    assert lib.msat_reset(plan, None, 1, None, None, None, None, 4, None) == _lib.MSAT_EINVAL
    assert lib.msat_step(plan, None, 1, None, None, None, 0, None, None, None, None, 0, None, 0, None, None, None,
                         None, 4, None) == _lib.MSAT_EINVAL
    # K outside 1..64, missing rng, both kinds of policy input at once
    dummy = C.addressof((C.c_char * 256)())
    a16 = (dummy + 127) // 128 * 128
    base = [plan, a16, 1, a16, a16, a16]
    tail = [None, 0, None, 0, None, None, None, None, 4, None]
    assert lib.msat_rollout_steps(*base, 0, a16, a16 + 64, 4, 0, None, None, None, 0, *tail) == _lib.MSAT_EINVAL
    assert lib.msat_rollout_steps(*base, 65, a16, a16 + 64, 4, 0, None, None, None, 0, *tail) == _lib.MSAT_EINVAL
    assert lib.msat_rollout_steps(*base, 2, None, a16 + 64, 4, 0, None, None, None, 0, *tail) == _lib.MSAT_EINVAL
    assert lib.msat_rollout_steps(*base, 2, a16, a16 + 64, 4, 0, a16, a16, None, 0, *tail) == _lib.MSAT_EINVAL
    # newly_satisfied needs the shaped reward; reward mode must be a known constant
    assert lib.msat_rollout_steps(*base, 2, a16, a16 + 64, 4, 0, None, None, None, 0, None, 0, None, 0, None, None,
                                  None, a16, 4, None) == _lib.MSAT_EINVAL
    assert lib.msat_plan_set_reward(plan, 7, 0.99, 0.02, 1.0) == _lib.MSAT_EINVAL
    assert lib.msat_tune(b"no_such_knob", 1) == _lib.MSAT_EINVAL
    h = C.c_void_p()
    assert lib.msat_host_pipe_create(C.byref(h), 0) == _lib.MSAT_EINVAL
    assert lib.msat_host_pipe_create(C.byref(h), 9) == _lib.MSAT_EINVAL
    assert lib.msat_host_wait(None, 0) == _lib.MSAT_EINVAL
    assert lib.msat_env_keys(None, None, 8, 4, 8, 3, None, None, None) == _lib.MSAT_EINVAL   # shard exceeds batch
    assert lib.msat_gae(None, 1, 1, None, None, None, 0.9, 0.9, None, None, None, 4, 4, None) == _lib.MSAT_EINVAL
    buf = (C.c_char * 4096)()
    addr = C.addressof(buf)
    mis = addr + 4 if (addr + 4) % 128 else addr + 8
    assert lib.msat_compile_bank(plan, addr, 1, mis, None) == _lib.MSAT_EALIGN


def test_no_cpu_fallback():
    env = M.SATEnv(20, 91, 8, verbose=False, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env.reset(np.ones((91, 3), np.int32), np.zeros(2, np.uint32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.calculate_gae(torch.zeros(2, 2), torch.zeros(2, 2, dtype=torch.bool), torch.zeros(2, 2), torch.zeros(2), 0.9, 0.9)


def test_shard_range_partitions_the_batch():
    for B in (1, 7, 16, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [M.shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == B
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1


def test_key_tensor_round_trip():
    keys = np.array([[0, 1], [0xFFFFFFFF, 0x80000000], [123456789, 4000000000]], np.uint32)
    t = M.env.as_u32_tensor(keys, torch.device("cpu"))
    assert t.dtype == torch.int32 and np.array_equal(M.env.u32_to_numpy(t), keys)
    t64 = M.env.as_u32_tensor(torch.from_numpy(keys.astype(np.int64)), torch.device("cpu"))
    assert np.array_equal(M.env.u32_to_numpy(t64), keys)


def test_config_reader_mirrors_runner_mapping(tmp_path):
    from marl_sat_b200 import config
    p = tmp_path / "cfg.yaml"
    p.write_text("SEED: 42\nenvironment:\n  NUM_VARS: 35\n  NUM_CLAUSES: 149\n  MAX_STEPS: 512\n  VARS_PER_AGENT: 7\n"
                 "  action_mode: 0\n  rewards:\n    R_CLAUSE: 0.0\n    R_SAT: 20.0\ntraining:\n  GAMMA: 0.995\n  NUM_ENVS: 128\n")
    cfg = config.load_config(str(p))
    flat = config.flatten(cfg)
    assert flat["NUM_VARS"] == 35 and flat["NUM_ENVS"] == 128
    env = config.make_env(cfg, verbose=False, device="cpu")
    assert env.num_agents == 5 and env.max_steps == 512 and env.r_sat == 20.0 and env.gamma == 0.995


def test_rollout_step_argument_validation():
    """The fused rollout step rejects overlapping rng buffers, shards outside the global batch and bad
    column counts before anything is enqueued (header contract)."""
    lib = _lib.load()
    env = M.SATEnv(20, 91, 8, verbose=False, device="cpu")
    plan = env._plan_for(3).handle
    buf = (C.c_uint32 * 64)()
    base = C.addressof(buf)
    a128 = (base + 127) & ~127                       # a 128-byte aligned fake "device" address inside buf
    rng, chain = a128, a128 + 4                      # overlapping: chain_out starts inside rng_in
    call = lambda rng_in, chain_out, Bg, off, B, rcols=1, dcols=1: lib.msat_rollout_step(
        plan, a128, 1, a128, a128, a128, rng_in, chain_out, Bg, off, None, a128, rcols, a128, dcols, None, None, None,
        B, None)
    assert call(rng, chain, 16, 0, 4) == _lib.MSAT_EINVAL
    assert call(a128, a128 + 64, 16, 14, 4) == _lib.MSAT_EINVAL          # shard [14, 18) exceeds the batch of 16
    assert call(a128, a128 + 64, 16, 0, 4, rcols=0) == _lib.MSAT_EINVAL
    assert call(a128, a128 + 64, 16, 0, 4, dcols=0) == _lib.MSAT_EINVAL
    assert call(None, a128 + 64, 16, 0, 4) == _lib.MSAT_EINVAL
    assert lib.msat_rollout_step_gnn(plan, a128, 1, a128, a128, a128, rng, chain, 16, 0, None, None, None, 0, None, 0,
                                     None, None, None, 4, None) == _lib.MSAT_EINVAL
    assert lib.msat_flip_gains(plan, None, 1, None, 4, 0.0, None, None, None) == _lib.MSAT_EINVAL
    assert lib.msat_gnn_dynamic(plan, None, 1, None, 4, None, None, None) == _lib.MSAT_EINVAL
    assert lib.msat_rollout_metrics(None, 1, 1, None, None, None, None, 2, 2, None, None) == _lib.MSAT_EINVAL


def test_dimacs_native_argument_validation():
    lib = _lib.load()
    rows, width = C.c_int32(), C.c_int32()
    assert lib.msat_dimacs_parse(None, 0, 0, None, None, C.byref(rows), C.byref(width), None, 0) == _lib.MSAT_EINVAL
    text = b"p cnf 3 1\n1 -2 3 0\n"
    assert lib.msat_dimacs_parse(text, len(text), 0, None, None, C.byref(rows), C.byref(width), None, 0) == 0
    assert (rows.value, width.value) == (1, 3)
    out = (C.c_int32 * 2)()
    assert lib.msat_dimacs_parse(text, len(text), 0, None, None, C.byref(rows), C.byref(width), out, 2) == _lib.MSAT_EINVAL
    bad = b"p cnf\n1 2 0\n"
    assert lib.msat_dimacs_parse(bad, len(bad), 0, None, None, C.byref(rows), C.byref(width), None, 0) == _lib.MSAT_EINVAL


def test_plan_groups_for_baseline_shapes():
    """Group sizes chosen by msat_plan_create for the BASELINE shapes (profiles/r1_group_size_sweep.md,
    profiles/r2_group_size_sweep.md): half-warp groups (two envs per warp) for the smallest observations."""
    expect = {(20, 91, None): 16, (50, 218, None): 32, (100, 430, None): 256, (250, 1065, None): 256,
              (100, 430, 7): 256, (35, 149, 7): 16}
    for (n, m, vpa), gs in expect.items():
        env = M.SATEnv(n, m, 512, vars_per_agent=vpa, verbose=False, device="cpu")
        assert env._plan_for(3).dims.group_threads == gs, (n, m, vpa)
    forced = M.SATEnv(100, 430, 512, verbose=False, device="cpu", group_threads=64)
    assert forced._plan_for(3).dims.group_threads == 64


def test_xla_ffi_shim_parses_and_covers_the_enqueue_entry_points():
    """csrc/xla_ffi_shim.cc cannot be built here (no JAX / XLA headers): it is syntax- and type-checked against
    the msat_* prototypes with a stand-in FFI header, and must bind every device enqueue entry point."""
    import re
    import shutil
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    shim = root / "marl_sat_b200" / "csrc" / "xla_ffi_shim.cc"
    text = shim.read_text()
    gxx = shutil.which("g++")
    if gxx:
        cuda_inc = next((p for p in ("/usr/local/cuda/include", "/usr/local/cuda-12.9/include") if Path(p).exists()), None)
        if cuda_inc:
            res = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-I", str(root / "tests" / "ffi_stub"), "-I", cuda_inc,
                                  str(shim)], capture_output=True, text=True)
            assert res.returncode == 0, res.stderr
    header = (root / "include" / "marl_sat_b200.h").read_text()
    declared = set(re.findall(r"\bint (msat_\w+)\(", header))
    host_only = {"msat_plan_create", "msat_plan_dims", "msat_plan_set_reward", "msat_plan_set_clause_update", "msat_plan_set_obs_dtype",
                 "msat_plan_set_reset_counter", "msat_tune", "msat_dimacs_parse", "msat_shutdown", "msat_host_pipe_create",
                 "msat_host_wait", "msat_rollout_step_host", "msat_rollout_step_host_async"}
    # single-step entry points are the K = 1 case of msat_rollout_steps, which the shim binds
    covered_by = {"msat_rollout_step": "msat_rollout_steps", "msat_rollout_step_gnn": "msat_rollout_steps"}
    for fn in sorted(declared - host_only):
        assert covered_by.get(fn, fn) + "(" in text, f"{fn} has no FFI handler"


def test_obs_dtype_option_and_its_group_sizes():
    """MSAT_OBS_INT8 is opt-in, validated, and re-tunes the group size (a quarter of the store trips)."""
    import pytest
    import torch
    with pytest.raises(ValueError):
        M.SATEnv(20, 91, 8, verbose=False, device="cpu", obs_dtype=torch.float32)
    expect = {(100, 430, None, 3): (256, 32), (250, 1065, None, 3): (256, 128), (100, 430, 7, 7): (256, 64),
              (20, 91, None, 3): (16, 16)}          # (n, m, vars per agent, literals per clause)
    for (n, m, vpa, k), (gs32, gs8) in expect.items():
        e32 = M.SATEnv(n, m, 512, vars_per_agent=vpa, verbose=False, device="cpu")
        e8 = M.SATEnv(n, m, 512, vars_per_agent=vpa, verbose=False, device="cpu", obs_dtype="int8")
        assert e32.obs_dtype == torch.int32 and e8.obs_dtype == torch.int8
        assert e32._plan_for(k).dims.group_threads == gs32 and e8._plan_for(k).dims.group_threads == gs8, (n, m, vpa)
    pinned = M.SATEnv(100, 430, 512, verbose=False, device="cpu", obs_dtype="int8", group_threads=128)
    assert pinned._plan_for(3).dims.group_threads == 128
    lib = _lib.load()
    assert lib.msat_plan_set_obs_dtype(None, 0) == _lib.MSAT_EINVAL
    assert lib.msat_plan_set_obs_dtype(pinned._plan_for(3).handle, 7) == _lib.MSAT_EINVAL
