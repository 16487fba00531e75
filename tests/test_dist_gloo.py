"""CPU tier: the N>1 path (contiguous env slices + statistics all-reduce) over gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_d2d_b200 import _lib
from gym_d2d_b200.dist import all_reduce_stats, shard_range, summarise


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_envs, out_q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        first, count = shard_range(total_envs, rank, world)
        # stand-in for the device stats vector of this rank's slice: every env contributes reward = its
        # GLOBAL index, so the reduced sums are independent of how the batch was cut
        idx = torch.arange(first, first + count, dtype=torch.float64)
        local = torch.zeros(_lib.NUM_STATS, dtype=torch.float64)
        local[0], local[1], local[2], local[3] = idx.sum(), 2 * idx.sum(), (idx * idx).sum(), count
        red = all_reduce_stats(local.clone())
        out_q.put((rank, first, count, red.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_stats_allreduce():
    world, total = 2, 1001
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    idx = torch.arange(total, dtype=torch.float64)
    expect = [idx.sum().item(), 2 * idx.sum().item(), (idx * idx).sum().item(), float(total)]
    assert res[0][1] == 0 and res[0][1] + res[0][2] == res[1][1] and res[1][1] + res[1][2] == total
    for _rank, _first, _count, red in res:
        assert red[:4] == pytest.approx(expect)
    s = summarise(torch.tensor(res[0][3], dtype=torch.float64))
    assert s['mean_reward'] == pytest.approx((total - 1) / 2) and s['env_steps'] == total


def test_all_reduce_is_noop_without_process_group():
    v = torch.arange(_lib.NUM_STATS, dtype=torch.float64)
    assert torch.equal(all_reduce_stats(v.clone()), v)
