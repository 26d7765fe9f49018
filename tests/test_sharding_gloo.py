"""world_size-2 (and 3) gloo tests on CPU for the N > 1 path: environment sharding and the
stats gather / max-over-ranks used by bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from die_b200.sharding import shard_range, gather_stats, max_over_ranks


def test_shard_ranges_partition_the_batch():
    for n, w in [(4096, 8), (10, 3), (5, 8), (7, 2), (1, 1)]:
        ranges = [shard_range(n, w, r) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_envs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard_range(n_envs, world, rank)
        # each environment's "reward" is a function of its global index only
        reward = torch.arange(a, b, dtype=torch.float64) * 1.5 - 3.0
        alive = torch.arange(a, b, dtype=torch.int64) * 7 + 1
        gr, ga = gather_stats(reward, alive, n_envs)
        t = max_over_ranks(10.0 + rank)
        q.put((rank, gr.tolist(), ga.tolist(), t))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_envs", [(2, 8), (2, 7), (3, 10)])
def test_gather_stats_and_max_over_ranks_gloo(world, n_envs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_r = (torch.arange(n_envs, dtype=torch.float64) * 1.5 - 3.0).tolist()
    exp_a = (torch.arange(n_envs, dtype=torch.int64) * 7 + 1).tolist()
    for rank, gr, ga, t in results:
        assert gr == exp_r and ga == exp_a
        assert t == 10.0 + world - 1


def test_single_process_passthrough():
    r, a = gather_stats(torch.tensor([1.0, 2.0], dtype=torch.float64), torch.tensor([3, 4]), 2)
    assert r.tolist() == [1.0, 2.0] and a.tolist() == [3, 4]
    assert max_over_ranks(2.5) == 2.5
