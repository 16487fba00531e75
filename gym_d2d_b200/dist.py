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


def bind_to_gpu_numa_node(device_index: int) -> Dict[str, object]:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off and prefer that node for new pages, so that the
    pinned host buffers of the end-to-end path (d2d_host_slot_buffers, torch pin_memory) are allocated next to the GPU's
    PCIe root instead of wherever rank 0's first touch put them.  Call BEFORE allocating pinned memory.  Best effort: returns
    what it found / did ({'node': -1, ...} when the platform exposes no topology - containers with one visible node)."""
    import ctypes
    import os
    info: Dict[str, object] = {'node': -1, 'cpus': None, 'mempolicy': False}
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        info['pci'] = bdf
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as f:
            node = int(f.read().strip())
        info['node'] = node
        if node < 0:
            return info
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info['cpus'] = len(allowed)
        # set_mempolicy(MPOL_PREFERRED, nodemask): syscall 238 on x86-64, 237 on aarch64
        import platform
        nr = {'x86_64': 238, 'aarch64': 237}.get(platform.machine())
        if nr is not None and node < 1024:
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            rc = ctypes.CDLL(None, use_errno=True).syscall(nr, 1, ctypes.byref(mask), 1024)
            info['mempolicy'] = rc == 0
    except Exception as exc:  # noqa: BLE001 - topology files are absent in many containers
        info['error'] = f'{type(exc).__name__}: {exc}'
    return info


class EpisodeStatsReducer:
    """Per-episode all-reduce of the statistics vector, off the step stream (BASELINE config #5).

    Two statistics buffers alternate: while episode k + 1 accumulates into one, a side stream sums episode k's replicas and
    all-reduces the [NUM_STATS] vector over the ranks (NCCL over NVLink on GPUs).  Usage per episode:

        red.begin_episode()          # binds + zeroes this episode's buffer on the step stream
        env.episode(...) / steps
        red.end_episode()            # side stream: wait for the episode, reduce, all-reduce; returns at once

    `results` holds one reduced float64 [NUM_STATS] device tensor per finished episode (read them after `finish()`)."""

    def __init__(self, env, group: Optional[dist.ProcessGroup] = None, keep: int = 64) -> None:
        self.env = env
        self.group = group
        self.bufs = [torch.zeros_like(env._stats), torch.zeros_like(env._stats)]
        self.free_events = [None, None]              # side stream done with buffer i
        self.side = torch.cuda.Stream(device=env.device)
        self.k = 0
        self.keep = keep
        self.results = []
        self.all_reduces = 0

    def begin_episode(self) -> None:
        i = self.k & 1
        cur = torch.cuda.current_stream(self.env.device)
        if self.free_events[i] is not None:
            cur.wait_event(self.free_events[i])      # the reduction of episode k - 2 has read this buffer
        self.env.bind_stats(self.bufs[i])
        self.env.reset_stats()

    def end_episode(self) -> None:
        i = self.k & 1
        cur = torch.cuda.current_stream(self.env.device)
        done = torch.cuda.Event()
        done.record(cur)
        self.side.wait_event(done)
        with torch.cuda.stream(self.side):
            v = self.bufs[i].sum(dim=0)
            all_reduce_stats(v, self.group)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                self.all_reduces += 1
            free = torch.cuda.Event()
            free.record(self.side)
        self.free_events[i] = free
        self.results.append(v)
        if len(self.results) > self.keep:
            self.results.pop(0)
        self.k += 1

    def finish(self) -> None:
        torch.cuda.current_stream(self.env.device).wait_stream(self.side)
