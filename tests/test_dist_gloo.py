"""world_size-2 tests on CPU (gloo) of the multi-rank host logic: contiguous env sharding, the
shard-invariant per-env key derivation (checked with the oracle's PRNG on each rank's slice) and the
product's own collectives -- ``allreduce_stats`` / ``mean_std_from_stats`` / ``metrics_from_sums``, the
functions ``normalize_advantages`` and ``rollout_metrics`` call (learner:530-532, 661-686; SURVEY.md section
8e).  The kernels around them run in tests/test_dist_gpu.py (two ranks on one GPU) and in bench.py's
multi-GPU self-check."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gae as ogae
from oracle import threefry as otf


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, Bg, P, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import marl_sat_b200 as M
        off, cnt = M.shard_range(Bg, world, rank)
        # per-env keys from GLOBAL indices: this rank's slice of the single-device arrays
        prob_key, reset_key = otf.prng_key(11), otf.prng_key(12)
        idx = otf.randint(prob_key, Bg, 0, P)[off:off + cnt]
        keys = otf.split(reset_key, Bg)[off:off + cnt]
        # advantage statistics: local (count, sum, sumsq) in float64, all-reduced like normalize_advantages()
        rng = np.random.default_rng(3)
        adv_global = rng.standard_normal((T, Bg)).astype(np.float32)
        adv = adv_global[:, off:off + cnt]
        local = torch.tensor([adv.size, adv.astype(np.float64).sum(), (adv.astype(np.float64) ** 2).sum()],
                             dtype=torch.float64)
        # the product's own collective + finalisation (what normalize_advantages / rollout_metrics call)
        stats = M.allreduce_stats(local)
        assert stats is not local and float(local[0]) == adv.size          # the local statistics stay local
        mean, std = M.mean_std_from_stats(stats)
        norm = (adv - np.float32(mean)) / (np.float32(std) + np.float32(1e-8))
        # rollout-metric sums: this rank's share of {reward, finished, solved, unsat at finish, steps of solved}
        done = rng.random((T, Bg)) < 0.3
        solved = done & (rng.random((T, Bg)) < 0.5)
        nunsat = rng.integers(0, 9, size=(T, Bg))
        estep = rng.integers(1, 50, size=(T, Bg))
        sl = np.s_[:, off:off + cnt]
        sums = torch.tensor([solved[sl].sum(), done[sl].sum(), (solved & done)[sl].sum(), (nunsat * done)[sl].sum(),
                             (estep * (solved & done))[sl].sum()], dtype=torch.float64)
        met = M.metrics_from_sums(M.allreduce_stats(sums), Bg)
        exp = {"mean_episodic_return": solved.sum() / Bg, "solve_rate": (solved & done).sum() / max(done.sum(), 1),
               "avg_unsatisfied_clauses": (nunsat * done).sum() / max(done.sum(), 1),
               "avg_steps_to_solve": (estep * (solved & done)).sum() / max((solved & done).sum(), 1)}
        for k_, v in exp.items():
            assert abs(met[k_] - float(v)) < 1e-9, (k_, met[k_], v)
        gathered = [None] * world
        dist.all_gather_object(gathered, (off, cnt, idx, keys, norm))
        if rank == 0:
            q.put(gathered)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("Bg", [16, 37])
def test_two_rank_sharding_and_advantage_allreduce(Bg):
    world, P, T = 2, 9, 5
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Bg, P, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    offs = [g[0] for g in gathered]
    assert offs[0] == 0 and offs[1] == gathered[0][1] and sum(g[1] for g in gathered) == Bg
    idx = np.concatenate([g[2] for g in gathered])
    keys = np.concatenate([g[3] for g in gathered])
    assert np.array_equal(idx, otf.randint(otf.prng_key(11), Bg, 0, P))
    assert np.array_equal(keys, otf.split(otf.prng_key(12), Bg))
    norm = np.concatenate([g[4] for g in gathered], axis=1)
    adv_global = np.random.default_rng(3).standard_normal((T, Bg)).astype(np.float32)
    assert np.allclose(norm, ogae.normalize_advantages(adv_global), rtol=1e-5, atol=1e-6)
