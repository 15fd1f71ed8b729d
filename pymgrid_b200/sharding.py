"""Multi-GPU sharding of a batch of microgrids: one process per GPU, contiguous env slices, no collective on the
step path (envs are independent: SURVEY.md section 8e).  The only collective is the optional aggregate used for
logging (total reward per step / per rollout), an all-reduce over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np



def shard_range(n_envs, rank, world):
    """Contiguous slice [lo, hi) of the global env ids owned by `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(n_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_env_config(global_env_config, rank, world):
    """The config index of every env of this rank's slice, and the global ids of those envs."""
    global_env_config = np.asarray(global_env_config)
    lo, hi = shard_range(len(global_env_config), rank, world)
    return global_env_config[lo:hi], np.arange(lo, hi)


def pymgrid25_env_config(global_batch):
    """BASELINE configs[2] / [4]: env i -> scenario i mod 25 (global numbering, independent of the sharding)."""
    return np.arange(global_batch) % 25


def aggregate_sum(x, group=None):
    """Sum of a per-rank tensor over all ranks (in place, returns x).  No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x


def sharded_pymgrid25(global_batch, rank, world, device=None, **kw):
    """This rank's engine for a global batch tiled over pymgrid25 (env i -> scenario i mod 25)."""
    from .engine import BatchedMicrogrid
    from .scenario import load_pymgrid25
    env_config, ids = shard_env_config(pymgrid25_env_config(global_batch), rank, world)
    bm = BatchedMicrogrid([load_pymgrid25(n) for n in range(25)], env_config, device=device, **kw)
    bm.global_env_ids = ids
    return bm


def total_reward(local_total):
    """Whole-job aggregate reward for logging: `local_total` is this rank's accumulator filled by the kernel
    (`BatchedMicrogrid.step(..., reward_total=t)` / `rollout(..., reward_total=t[n_steps])`: warp-shuffle reduction + one
    atomicAdd per warp); the only collective is this all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    return aggregate_sum(local_total)
