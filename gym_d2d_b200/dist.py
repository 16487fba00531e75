"""Multi-GPU plumbing: one process per GPU, each owning a contiguous slice of the global env batch.

Environments are independent (envs/d2d_env.py:29,41-43: every D2DEnv owns all of its state), so the step
path needs NO collective.  The only exchange is the episode-statistics vector (sum reward, sum capacity,
sum reward^2, env-steps, penalties, rescues), all-reduced per episode / log interval over
torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [first, first + count) of the global batch owned by `rank`; the first
    total_envs % world_size ranks hold one extra env."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('bad rank / world_size')
    base, extra = divmod(int(total_envs), world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def all_reduce_stats(local_stats: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum the [NUM_STATS] float64 statistics vector over ranks (in place) and return it.
    No-op when torch.distributed is not initialised (single GPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_stats, op=dist.ReduceOp.SUM, group=group)
    return local_stats


def summarise(stats: torch.Tensor) -> Dict[str, float]:
    """Derived episode statistics from a (globally reduced) stats vector."""
    v = dict(zip(_lib.STAT_NAMES, stats.detach().cpu().tolist()))
    n = max(v['env_steps'], 1.0)
    mean = v['sum_reward'] / n
    v['mean_reward'] = mean
    v['var_reward'] = max(v['sum_reward_sq'] / n - mean * mean, 0.0)
    v['mean_capacity_mbps'] = v['sum_capacity_mbps'] / n
    return v


def make_sharded_env(total_envs: int, env_config: Optional[dict] = None, seed: int = 0, **kwargs):
    """VecD2DEnv over this rank's slice (rank / world size / device from torch.distributed + LOCAL_RANK)."""
    import os
    from .vec_env import VecD2DEnv
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    first, count = shard_range(total_envs, rank, world)
    local = int(os.environ.get('LOCAL_RANK', rank % max(torch.cuda.device_count(), 1)))
    return VecD2DEnv(count, env_config, device=torch.device('cuda', local), seed=seed, global_env_offset=first,
                     **kwargs)
