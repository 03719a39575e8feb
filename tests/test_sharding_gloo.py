"""world_size-2 gloo test of the N>1 host logic (shard assignment + result aggregation)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from erweitern_55_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.shard_range(1001, rank, world)
    dist.barrier()
    total, secs = sharding.aggregate(b - a, 0.5 + rank)
    got = sharding.exchange("moves", {"moves": 10 * (rank + 1), "seconds": 1.5 - rank, "error": None if rank else "x"})
    again = sharding.exchange("again", rank)
    q.put((rank, a, b, total, secs, got, again))
    dist.destroy_process_group()


def test_two_rank_aggregation():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 501), (501, 1001)]
    assert all(r[3] == 1001.0 and r[4] == 1.5 for r in res)
    # the host-side meeting point bench.py uses around the GPU-resident self-play: every rank sees every rank's payload
    want = [{"moves": 10, "seconds": 1.5, "error": "x"}, {"moves": 20, "seconds": 0.5, "error": None}]
    assert all(r[5] == want and r[6] == [0, 1] for r in res)
    assert sharding.exchange("alone", {"a": 1}) == [{"a": 1}]           # not distributed: identity
