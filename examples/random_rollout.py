#!/usr/bin/env python
"""Random-policy rollout through the drop-in API (the role of the reference's
``src/runners/no_policy.py`` smoke script): load DIMACS formulas (or synthesise uniform 3-SAT), roll out T
steps of B auto-resetting envs with uniformly random actions, compute GAE with a zero critic and print the
reference's rollout metrics.

    python examples/random_rollout.py --cnf-dir tests/golden --num-vars 20 --num-clauses 91 --envs 1024
    python examples/random_rollout.py --num-vars 100 --num-clauses 430 --envs 65536 --steps 128
"""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import marl_sat_b200 as M  # noqa: E402
from marl_sat_b200 import dimacs  # noqa: E402
from marl_sat_b200.synth import uniform_ksat  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cnf-dir", default=None)
    ap.add_argument("--num-vars", type=int, default=20)
    ap.add_argument("--num-clauses", type=int, default=91)
    ap.add_argument("--vars-per-agent", type=int, default=None)
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--max-steps", type=int, default=512)
    ap.add_argument("--seed", type=int, default=42)
    a = ap.parse_args()

    if a.cnf_dir:
        probs = [p for p in dimacs.load_cnf_problems(a.cnf_dir)         # Python reader (reference semantics)
                 if p["num_vars"] == a.num_vars and p["num_clauses"] == a.num_clauses]
        clauses = dimacs.stack_problems(probs)                          # dimacs.load_cnf_bank_array = native reader
    else:
        clauses = uniform_ksat(256, a.num_vars, a.num_clauses, 3, seed=a.seed)
    env = M.SATEnv(a.num_vars, a.num_clauses, a.max_steps, vars_per_agent=a.vars_per_agent)
    vec = M.VecSATEnv(env, clauses, a.envs, M.prng_key(a.seed), compact_outputs=True)
    obs = vec.reset()
    buf = M.RolloutBuffer(env, vec.bank, a.steps, a.envs)
    gen = torch.Generator(device=obs.device).manual_seed(a.seed)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(a.steps):
        actions = torch.randint(0, env.max_vars_per_agent + 1, (a.envs, env.num_agents), generator=gen,
                                device=obs.device, dtype=torch.int32)
        buf.state[t].copy_(vec.state)
        buf.action[t].copy_(actions)
        vec.step(actions, out=buf.step_outputs(t, vec.out["obs"]))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = torch.zeros(3, dtype=torch.float64, device=obs.device)
    adv, targets = M.calculate_gae(buf.reward, buf.global_done, buf.value, torch.zeros(a.envs, device=obs.device),
                                   0.995, 0.95, stats=stats)
    M.normalize_advantages(adv, stats=stats)
    metrics = M.rollout_metrics(buf.reward, buf.global_done, buf.solved, buf.num_unsatisfied, buf.episode_step)
    print(f"{a.envs} envs x {a.steps} steps of {env.num_agents} agents: {a.envs * a.steps / dt / 1e6:.2f} M env-steps/s "
          f"(incl. torch.randint actions)")
    print({k: round(v, 4) for k, v in metrics.items()}, "adv mean/std after normalisation:",
          round(float(adv.mean()), 5), round(float(adv.std(unbiased=False)), 5))


if __name__ == "__main__":
    main()
