"""Launches the observation-writing step kernel at a small shape so that one `ncu` run can capture it:

    ncu --set full --import-source on --clock-control none -k regex:env_kernel -c 6 -o gpurun_out/small \
        python profiles/capture_small.py uf20-91 65536

(run only after the plain `python profiles/capture_small.py ...` has exited 0).  Not a benchmark."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import marl_sat_b200 as M                       # noqa: E402
from marl_sat_b200 import synth                  # noqa: E402

SHAPES = {"uf20-91": (20, 91), "uf35-149": (35, 149), "uf50-218": (50, 218), "uf100-430": (100, 430), "uf250-1065": (250, 1065)}
n, m = SHAPES[sys.argv[1]]
B = int(sys.argv[2])
GS = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
env = M.SATEnv(n, m, 512, verbose=False, device=dev, group_threads=GS)
bank = env.make_bank(synth.uniform_ksat_torch(B, n, m, 3, seed=1, device=dev), validate=False)
g = torch.Generator(device=dev).manual_seed(0)
acts = torch.randint(0, env.max_vars_per_agent + 1, (8, B, env.num_agents), generator=g, device=dev, dtype=torch.int32)
vec = M.VecSATEnv(env, bank, B, M.prng_key(1))
vec.reset()
for i in range(4):
    vec.step(acts[i])
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(20):
    vec.step(acts[i % 8])
ev1.record()
torch.cuda.synchronize()
print(f"{sys.argv[1]} x {B} GS={GS}: {ev0.elapsed_time(ev1) / 20 * 1e3:.1f} us/step")
