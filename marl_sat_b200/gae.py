"""MAPPO advantage path on device: GAE / return scan and global advantage normalisation.

Mirrors ``_calculate_gae`` and the normalisation that follows it in the reference learner
(``/root/reference/src/learners/mappo_gnn_sat_learner.py:504-532``).  With ``torch.distributed``
initialised, the three normalisation statistics (count, sum, sum of squares; float64) are
all-reduced so the result equals the single-device global mean/std over all ``T x B_global``
elements -- the only collective on the advantage path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .env import _ptr, _stream_ptr


def calculate_gae(reward: torch.Tensor, done: torch.Tensor, value: torch.Tensor, last_val: torch.Tensor,
                  gamma: float, gae_lambda: float, stats: Optional[torch.Tensor] = None,
                  out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """reward f32 ``[T,B,A]`` (agent 0 is read, learner:514) or ``[T,B]``; done bool/uint8 ``[T,B]``;
    value f32 ``[T,B]``; last_val f32 ``[B]`` -> ``(advantages, targets)`` f32 ``[T,B]``
    (``targets = advantages + value`` with un-normalised advantages, learner:526).  ``stats`` (a zeroed
    float64[3] device tensor) receives the advantages' (count, sum, sum of squares) from the same pass.
    Within one time segment the operation order is the reference's (no FMA contraction); when the batch is
    too small to fill the GPU the time axis is split into segments whose affine maps are composed, which
    changes the rounding sequence (still within the 1e-5 relative tolerance of the contract)."""
    lib = _lib.load()
    if not reward.is_cuda:
        raise RuntimeError("calculate_gae needs CUDA tensors: there is no CPU fallback")
    if value.dim() != 2:
        raise ValueError(f"value must be [T, B], got {tuple(value.shape)}")
    T, B = value.shape
    if reward.dtype != torch.float32 or value.dtype != torch.float32 or last_val.dtype != torch.float32:
        raise TypeError("reward/value/last_val must be float32")
    if reward.dim() not in (2, 3) or tuple(reward.shape[:2]) != (T, B):
        raise ValueError(f"reward must be [T, B] or [T, B, A] with T, B = {T}, {B}; got {tuple(reward.shape)}")
    if tuple(done.shape) != (T, B) or done.dtype not in (torch.bool, torch.uint8):
        raise ValueError(f"done must be bool/uint8 [T, B] = [{T}, {B}], got {done.dtype} {tuple(done.shape)}")
    if tuple(last_val.shape) != (B,):
        raise ValueError(f"last_val must be [B] = [{B}], got {tuple(last_val.shape)}")
    for name, t in (("done", done), ("value", value), ("last_val", last_val)):
        if t.device != reward.device:
            raise ValueError(f"{name} is on {t.device}, reward on {reward.device}")
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() != 3 or stats.device != reward.device
                              or not stats.is_contiguous()):
        raise ValueError("stats must be a contiguous float64[3] tensor on the same device (zero it before the call)")
    if done.dtype == torch.bool:
        done = done.view(torch.uint8)
    done = done.contiguous()
    value = value.contiguous()
    last_val = last_val.contiguous()
    rs_t, rs_b = reward.stride(0), reward.stride(1)
    if out is not None:         # preallocated (advantages, targets): contiguous float32 [T, B]
        adv, tgt = out
        for t in (adv, tgt):
            if tuple(t.shape) != (T, B) or t.dtype != torch.float32 or not t.is_contiguous() or t.device != value.device:
                raise ValueError("out must be two contiguous float32 [T, B] tensors on the inputs' device")
    else:
        adv = torch.empty((T, B), dtype=torch.float32, device=value.device)
        tgt = torch.empty((T, B), dtype=torch.float32, device=value.device)
    _lib.check(lib.msat_gae(_ptr(reward), rs_t, rs_b, _ptr(done), _ptr(value), _ptr(last_val), float(gamma),
                            float(gae_lambda), _ptr(adv), _ptr(tgt), _ptr(stats), T, B, _stream_ptr(value.device)),
               "msat_gae")
    return adv, tgt


def advantage_stats(adv: torch.Tensor, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Accumulate ``(count, sum, sum of squares)`` of ``adv`` into a float64[3] device tensor."""
    lib = _lib.load()
    if stats is None:
        stats = torch.zeros(3, dtype=torch.float64, device=adv.device)
    _lib.check(lib.msat_adv_stats(_ptr(adv), adv.numel(), _ptr(stats), _stream_ptr(adv.device)), "msat_adv_stats")
    return stats


def _world_size(group=None) -> int:
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size(group)
    return 1


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a rank-local float64 statistics vector -- (count, sum, sum of squares) of the advantages, or the
    five rollout-metric sums -- over all ranks.  Returns a private reduced copy (the caller's local values
    stay local); the input itself when ``torch.distributed`` is not initialised or has one rank.  Works for
    CPU (gloo) and CUDA (nccl) tensors: this is the only collective of the advantage path."""
    if _world_size(group) <= 1:
        return stats
    out = stats.clone()
    torch.distributed.all_reduce(out, op=torch.distributed.ReduceOp.SUM, group=group)
    return out


def mean_std_from_stats(stats: torch.Tensor):
    """``(mean, population std)`` from (count, sum, sum of squares), in float64 like ``adv_normalize_kernel``."""
    n, s, ss = (float(x) for x in stats.tolist())
    mean = s / n
    return mean, max(ss / n - mean * mean, 0.0) ** 0.5


def normalize_advantages(adv: torch.Tensor, stats: Optional[torch.Tensor] = None, group=None,
                         reduced: bool = False) -> torch.Tensor:
    """In place ``adv = (adv - mean) / (std + 1e-8)`` with the global population std (learner:530-532).
    ``stats``: the local (count, sum, sum of squares) if ``calculate_gae`` already produced them (saves a
    pass over ``adv``); with ``torch.distributed`` initialised they are all-reduced first (a private copy)."""
    lib = _lib.load()
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() != 3):
        raise ValueError("stats must be float64[3]")
    work = adv if adv.is_contiguous() else adv.contiguous()      # a strided view is normalised through a copy ...
    if stats is None:
        stats = advantage_stats(work)
    if not reduced:                                              # reduced=True: `stats` are already global
        stats = allreduce_stats(stats, group)                    # global (count, sum, sum of squares)
    _lib.check(lib.msat_adv_normalize(_ptr(work), work.numel(), _ptr(stats), _stream_ptr(work.device)),
               "msat_adv_normalize")
    if work is not adv:
        adv.copy_(work)                                          # ... and written back: the call is in place
    return adv
