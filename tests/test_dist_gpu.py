"""Two ranks on ONE GPU (gloo backend, CUDA tensors): the distributed branches of the product code --
``VecSATEnv`` shards, ``calculate_gae`` + ``normalize_advantages`` (all-reduce of the statistics),
``rollout_metrics`` (all-reduce of the sums) and the ``MAPPOTrainer`` update with DDP gradient all-reduce --
must reproduce the single-process full-batch results.  (NCCL refuses two ranks on one device, gloo
stages CUDA tensors through the host; the NCCL path itself is exercised by bench.py on 2-8 GPUs.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(M, world, rank, Bg=48, T=12):
    from marl_sat_b200.synth import uniform_ksat
    n, m = 20, 40
    env = M.SATEnv(n, m, 5, verbose=False)
    bank = env.make_bank(uniform_ksat(7, n, m, 3, seed=3))
    vec = M.VecSATEnv(env, bank, Bg, M.prng_key(11), world_size=world, rank=rank, compact_outputs=True)
    vec.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    actions = torch.randint(0, 5, (T, Bg, env.num_agents), generator=g, device="cuda", dtype=torch.int32)
    values = torch.randn((T + 1, Bg), generator=g, device="cuda")
    return env, bank, vec, actions, values


def _rollout_adv_metrics(M, env, vec, actions, values, T):
    sl = slice(vec.env_offset, vec.env_offset + vec.num_envs)
    buf = M.RolloutBuffer(env, vec.bank, T, vec.num_envs)
    buf.collect(vec, lambda t, v: (actions[t, sl], values[t, sl], None))
    stats = torch.zeros(3, dtype=torch.float64, device="cuda")
    adv, tgt = M.calculate_gae(buf.reward[:, :, 0], buf.global_done, buf.value, values[T, sl].contiguous(), 0.995, 0.95,
                               stats=stats)
    norm = M.normalize_advantages(adv.clone(), stats=stats)
    met = M.rollout_metrics(buf.reward, buf.global_done, buf.solved, buf.num_unsatisfied, buf.episode_step,
                            num_envs_global=vec.num_envs_global)
    return adv, norm, met


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import marl_sat_b200 as M
        T = 12
        env, bank, vec, actions, values = _setup(M, world, rank, T=T)
        adv, norm, met = _rollout_adv_metrics(M, env, vec, actions, values, T)
        out = [None] * world
        dist.all_gather_object(out, (vec.env_offset, adv.cpu().numpy(), norm.cpu().numpy(), met))
        if rank == 0:
            q.put(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_one_gpu_match_single_process():
    import marl_sat_b200 as M
    world, T = 2, 12
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    env, bank, vec, actions, values = _setup(M, 1, 0, T=T)
    adv, norm, met = _rollout_adv_metrics(M, env, vec, actions, values, T)
    gathered.sort(key=lambda g: g[0])
    assert np.array_equal(np.concatenate([g[1] for g in gathered], axis=1), adv.cpu().numpy())       # env sharding
    got = np.concatenate([g[2] for g in gathered], axis=1)
    assert np.allclose(got, norm.cpu().numpy(), rtol=1e-5, atol=1e-6)     # global mean/std through the all-reduce
    for g in gathered:
        for k_, v in met.items():
            assert abs(g[3][k_] - v) <= 1e-9 * max(1.0, abs(v)), k_
    assert met["solve_rate"] > 0 or met["avg_unsatisfied_clauses"] > 0


def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import marl_sat_b200 as M
        from marl_sat_b200.mappo import MAPPOTrainer, MLPActorCritic, PPOConfig
        env, bank, vec, _, _ = _setup(M, world, rank, Bg=32, T=8)
        torch.manual_seed(0)
        net = MLPActorCritic(env, hidden=32)
        tr = MAPPOTrainer(vec, net, PPOConfig(num_steps=8, update_epochs=2, minibatch_size=32), seed=1, autocast=False)
        m1 = tr.train_cycle()
        m2 = tr.train_cycle()
        flat = torch.cat([p.detach().reshape(-1) for p in tr.raw_net.parameters()]).cpu()
        out = [None] * world
        dist.all_gather_object(out, (flat.numpy(), m1, m2))
        if rank == 0:
            q.put(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_mappo_update_keeps_replicas_in_sync():
    """rollout -> GAE -> PPO update on two ranks: the gradient all-reduce keeps the replicas identical, the
    losses are finite and the all-reduced metrics agree on both ranks."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    (w0, a1, a2), (w1, b1, b2) = out
    assert np.array_equal(w0, w1) and np.isfinite(w0).all()
    for k_ in ("mean_episodic_return", "solve_rate", "avg_unsatisfied_clauses", "avg_steps_to_solve"):
        assert a2[k_] == b2[k_]
    assert np.isfinite([a2["value_loss"], a2["actor_loss"], a2["entropy"]]).all() and a2["optimizer_steps"] == 2 * (8 * 16 // 32)
