"""Manual stress run (not collected by pytest): hammer the time-out / auto-reset boundary at the headline
shape for every thread-group size.  A dead-lock shows up as the `timeout` of the calling shell.

    timeout 120 python tests/stress_boundary.py
"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import marl_sat_b200 as M  # noqa: E402
from marl_sat_b200.synth import uniform_ksat_torch  # noqa: E402


def main():
    dev = torch.device("cuda")
    for (n, m, B) in ((100, 430, 65536), (50, 218, 65536), (20, 91, 131072)):
        problems = uniform_ksat_torch(4096, n, m, 3, seed=1, device=dev)
        for gs in (0, 16, 32, 64, 128, 256):
            for max_steps in (2, 3, 5):
                env = M.SATEnv(n, m, max_steps, verbose=False, group_threads=gs)
                vec = M.VecSATEnv(env, problems, B, M.prng_key(gs + max_steps), emit_obs=(gs in (0, 16, 256)))
                vec.reset()
                g = torch.Generator(device=dev).manual_seed(0)
                acts = torch.randint(0, env.max_vars_per_agent + 1, (8, B, env.num_agents), generator=g, device=dev,
                                     dtype=torch.int32)
                t0 = time.perf_counter()
                for t in range(120):
                    out = vec.step(acts[t % 8])
                torch.cuda.synchronize()
                es = out["episode_step"]
                assert int(es.min()) >= 1 and int(es.max()) <= max_steps
                print(f"n={n} B={B} gs={gs or 'auto'} max_steps={max_steps}: 120 steps ok "
                      f"({time.perf_counter() - t0:.2f} s)", flush=True)
                del vec, env


if __name__ == "__main__":
    main()
