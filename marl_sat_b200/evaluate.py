"""Greedy evaluation loop (next-tier row): the env half of
``evaluate_policy`` (``/root/reference/src/runners/mappo_runner.py:30-73``).

``reset`` -> ``max_steps`` x (policy -> ``step_env``) **without** auto-reset (the env keeps stepping past
``done``, runner:50), then the first step at which the formula was solved, ``steps_to_solve``
(``max_steps`` when never solved) and the solving assignment (zeros when never solved).  The policy is a
callable ``policy_fn(obs int32[B,A,D], state) -> actions`` (greedy ``argmax`` of the logits in the
reference); the bookkeeping of runner:57-70 runs on device (``msat_eval_track``).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch

from . import _lib
from .env import FormulaBank, SATEnv, SATState, _ptr, _stream_ptr, as_u32_tensor


def evaluate_policy(policy_fn: Callable, env: SATEnv, bank: FormulaBank, problem_idx: torch.Tensor, keys,
                    max_steps: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns ``(was_ever_solved bool[B], steps_to_solve i32[B], solution_assignments i32[B,n])``."""
    lib = _lib.load()
    dev = env._require_cuda()
    B = int(problem_idx.shape[0])
    d = bank.plan.dims
    obs, state = env.reset_from_bank(bank, problem_idx, as_u32_tensor(keys, dev))
    out = env.alloc_step_outputs(B, d)
    out["obs"] = obs
    ever = torch.zeros((B,), dtype=torch.uint8, device=dev)
    steps = torch.full((B,), int(max_steps), dtype=torch.int32, device=dev)      # runner:67
    solution = torch.zeros((B, d.n), dtype=torch.int32, device=dev)              # runner:63
    packed = state.packed
    for t in range(max_steps):
        actions = policy_fn(out["obs"], SATState(env, bank, packed, True))
        actions = env._actions_tensor(actions, B, dev)
        env.step_into(bank, packed, packed, actions, out, auto_reset=False)
        _lib.check(lib.msat_eval_track(bank.plan.handle, _ptr(packed), _ptr(out["solved"]), t, B, _ptr(ever),
                                       _ptr(steps), _ptr(solution), _stream_ptr(dev)), "msat_eval_track")
    return ever.bool(), steps, solution
