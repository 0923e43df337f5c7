"""N>1 host logic on CPU: two gloo ranks shard rows, time a fake step and reduce the max."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from basic_dsp_b200 import sharding


def test_row_shard_partitions():
    for total in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            blocks = [sharding.row_shard(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = sharding.row_shard(64, world, rank)
    rows = torch.arange(start, stop)
    # every rank "processes" its own rows; only the elapsed time is reduced
    elapsed = 0.010 * (rank + 1)
    mx = sharding.max_over_ranks(elapsed)
    total = torch.tensor([float(rows.sum())])
    dist.all_reduce(total)  # test-only check that the shards cover 0..63 exactly once
    out.put((rank, start, stop, mx, float(total.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_sharding():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 32), (32, 64)]
    assert all(abs(r[3] - 0.020) < 1e-12 for r in res)          # max over ranks
    assert all(r[4] == sum(range(64)) for r in res)
    assert sharding.weak_scaling_throughput(100, 10, 0.5, 2) == 4000.0
