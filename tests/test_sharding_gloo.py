"""Multi-process (world_size 2, gloo, CPU) test of the N>1 host logic: contiguous sharding covers the global batch
exactly once, and the logging aggregate (all-reduce of per-rank reward sums) equals the single-process total.
The per-rank 'compute' here is the C oracle (test infrastructure) since there is no GPU in the CPU suite."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pymgrid_b200.sharding import pymgrid25_env_config, shard_env_config, shard_range


def test_shard_ranges_partition_the_batch():
    for n, world in ((65536, 8), (1048576, 8), (10, 3), (7, 8), (25, 1)):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, n_steps, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import OracleBatch
    from pymgrid_b200.scenario import load_pymgrid25
    from pymgrid_b200.sharding import aggregate_sum
    configs = [load_pymgrid25(n) for n in range(25)]
    env_config, ids = shard_env_config(pymgrid25_env_config(global_batch), rank, world)
    actions = np.random.default_rng(5).random((n_steps, global_batch, 4))[:, ids]    # same global stream, own slice
    rewards, _, _ = OracleBatch([configs[c] for c in env_config]).rollout(actions)
    per_step = torch.from_numpy(rewards.sum(axis=1))
    aggregate_sum(per_step)                                                          # the only collective
    covered = torch.zeros(global_batch, dtype=torch.int32)
    covered[torch.from_numpy(ids)] = 1
    dist.all_reduce(covered)
    if rank == 0:
        out_q.put((per_step.numpy(), covered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_aggregate_matches_single_process():
    from oracle.oracle import OracleBatch
    from pymgrid_b200.scenario import load_pymgrid25
    global_batch, n_steps, world = 101, 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, n_steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    per_step, covered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (covered == 1).all()
    configs = [load_pymgrid25(n) for n in range(25)]
    actions = np.random.default_rng(5).random((n_steps, global_batch, 4))
    rewards, _, _ = OracleBatch([configs[c] for c in pymgrid25_env_config(global_batch)]).rollout(actions)
    np.testing.assert_allclose(per_step, rewards.sum(axis=1), rtol=1e-12)


# ---- the composed path (any module list): same sharding, per-rank compute = the host build of its C source ------------
def _composed_worker(rank, world, port, global_batch, n_steps, out_q):
    import ctypes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pymgrid_b200.compose import ComposedBatch
    from pymgrid_b200.sharding import aggregate_sum
    from tests import hostsim
    from tests.compose_cases import load_cases
    case = next(c for c in load_cases() if c.label == "several_of_each")
    configs = [case.modules(), case.modules()]
    env_config, ids = shard_env_config(np.arange(global_batch) % 2, rank, world)
    hostsim.select(ctypes.CDLL(hostsim.build()))
    batch = ComposedBatch(configs, env_config, microgrid_kwargs=case.microgrid_kwargs)
    actions = np.random.default_rng(5).random((n_steps, global_batch, batch.comp.n_act))[:, ids]
    out = batch.rollout(torch.from_numpy(np.ascontiguousarray(actions)), obs=False)
    per_step = out["reward"].sum(dim=1)
    aggregate_sum(per_step)                                      # the only collective: the logging aggregate
    if rank == 0:
        out_q.put((per_step.numpy(), out["reward"].numpy(), ids))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_composed_batch_matches_single_process():
    import ctypes
    from pymgrid_b200.compose import ComposedBatch
    from tests import hostsim
    from tests.compose_cases import load_cases
    global_batch, n_steps, world = 37, 5, 2
    hostsim.build()                                              # once, before the ranks race to build it
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_composed_worker, args=(r, world, port, global_batch, n_steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    per_step, rank0_rewards, rank0_ids = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = next(c for c in load_cases() if c.label == "several_of_each")
    hostsim.select(ctypes.CDLL(hostsim.build()))
    batch = ComposedBatch([case.modules(), case.modules()], np.arange(global_batch) % 2, microgrid_kwargs=case.microgrid_kwargs)
    actions = np.random.default_rng(5).random((n_steps, global_batch, batch.comp.n_act))
    want = batch.rollout(torch.from_numpy(actions), obs=False)["reward"].numpy()
    assert np.array_equal(rank0_rewards, want[:, rank0_ids])     # a shard computes exactly its slice of the global batch
    np.testing.assert_allclose(per_step, want.sum(axis=1), rtol=1e-12)
