"""Launches each hot kernel a few times at the headline shape so that one `ncu` run can capture them:

    ncu --set full --import-source on --clock-control none -k regex:'env_kernel|gae|normalize' -c 24 \
        -o gpurun_out/r2_kernels python profiles/capture_kernels.py

(run only after the plain `python profiles/capture_kernels.py` has exited 0).  Not a benchmark."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import marl_sat_b200 as M                       # noqa: E402
from marl_sat_b200 import _lib, synth            # noqa: E402

which = set(sys.argv[1:]) or {"step", "gnn", "gae", "gae_plain", "norm", "multi", "incr"}
dev = torch.device("cuda", 0)
n, m, k, B, T = 100, 430, 3, 65536, 512
env = M.SATEnv(n, m, 512, verbose=False, device=dev)
bank = env.make_bank(synth.uniform_ksat_torch(B, n, m, k, seed=1, device=dev), validate=False)
g = torch.Generator(device=dev).manual_seed(0)
acts = torch.randint(0, 5, (8, B, env.num_agents), generator=g, device=dev, dtype=torch.int32)
if "step" in which:
    vec = M.VecSATEnv(env, bank, B, M.prng_key(1))
    vec.reset()
    for i in range(3):
        vec.step(acts[i])
if "gnn" in which:
    vg = M.VecSATEnv(env, bank, B, M.prng_key(2), emit_obs=False, compact_outputs=True, gnn_outputs=True)
    vg.reset()
    for i in range(3):
        vg.step(acts[i])
if "multi" in which:
    env2 = M.SATEnv(50, 218, 512, verbose=False, device=dev)
    bank2 = env2.make_bank(synth.uniform_ksat_torch(4096, 50, 218, 3, seed=1, device=dev), validate=False)
    v2 = M.VecSATEnv(env2, bank2, 4096, M.prng_key(3))
    v2.reset()
    a2 = torch.randint(0, 9, (32, 4096, env2.num_agents), generator=g, device=dev, dtype=torch.int32)
    out = v2.alloc_multi_step_outputs(32, emit_every_step=True)
    for i in range(2):
        v2.steps(a2, out)
if which & {"gae", "gae_plain", "norm"}:
    reward = (torch.rand((T, B), generator=g, device=dev) < 0.01).float()
    done = (torch.rand((T, B), generator=g, device=dev) < 0.005).to(torch.uint8)
    value = torch.randn((T, B), generator=g, device=dev)
    last = torch.randn((B,), generator=g, device=dev)
    stats = torch.zeros(3, dtype=torch.float64, device=dev)
    adv = None
    if "gae" in which:
        for i in range(2):
            stats.zero_()
            adv, tgt = M.calculate_gae(reward, done, value, last, 0.995, 0.95, stats=stats)
    if "gae_plain" in which:
        _lib.load().msat_tune(b"gae_plain", 1)
        for i in range(2):
            stats.zero_()
            adv, tgt = M.calculate_gae(reward, done, value, last, 0.995, 0.95, stats=stats)
        _lib.load().msat_tune(b"gae_plain", 0)
    if "norm" in which and adv is not None:
        for i in range(2):
            M.normalize_advantages(adv, stats=stats)
if "incr" in which:
    env_i = M.SATEnv(n, m, 512, verbose=False, device=dev, clause_update="incremental")
    bank_i = env_i.make_bank(synth.uniform_ksat_torch(B, n, m, k, seed=1, device=dev), validate=False)
    vi = M.VecSATEnv(env_i, bank_i, B, M.prng_key(2), emit_obs=False, compact_outputs=True, gnn_outputs=True)
    vi.reset()
    for i in range(3):
        vi.step(acts[i])
    del vi, bank_i
if "i8" in which:
    env8 = M.SATEnv(n, m, 512, verbose=False, device=dev, obs_dtype=torch.int8)
    v8 = M.VecSATEnv(env8, bank.for_env(env8), B, M.prng_key(1))
    v8.reset()
    for i in range(3):
        v8.step(acts[i])
    del v8
if "shapes" in which:
    # one observation-writing step launch per (workload, per-GPU batch) for profiles/traffic.json
    SHAPES = [("uf100-430", 100, 430, 3, None, "uniform", bs) for bs in (32768, 16384, 8192)] + [
        ("uf50-218", 50, 218, 3, None, "uniform", 4096), ("uf250-1065", 250, 1065, 3, None, "uniform", 16384),
        ("uf250-1065", 250, 1065, 3, None, "uniform", 2048), ("mixed-k3-7", 100, 430, 7, 7, "mixed", 32768),
        ("mixed-k3-7", 100, 430, 7, 7, "mixed", 4096), ("uf20-91", 20, 91, 3, None, "uniform", 65536)]
    for name, n_, m_, k_, vpa, kind, bs in SHAPES:
        e_ = M.SATEnv(n_, m_, 512, vars_per_agent=vpa, verbose=False, device=dev)
        pr = (synth.mixed_ksat_torch(bs, n_, m_, 3, k_, seed=1, device=dev) if kind == "mixed"
              else synth.uniform_ksat_torch(bs, n_, m_, k_, seed=1, device=dev))
        v_ = M.VecSATEnv(e_, e_.make_bank(pr, validate=False), bs, M.prng_key(1))
        v_.reset()
        a_ = torch.randint(0, e_.max_vars_per_agent + 1, (bs, e_.num_agents), generator=g, device=dev, dtype=torch.int32)
        for i in range(2):
            v_.step(a_)
        print("shape", name, bs)
        del v_, pr
torch.cuda.synchronize()
print("ok")
