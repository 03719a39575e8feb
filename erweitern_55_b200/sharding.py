"""Sharding of independent evaluations / games across the GPUs of one node.

The reference scales only by running more independent ``client.py`` processes (client.py:44-70); here one
process per GPU owns one context and a contiguous share of the work.  There is no collective on the
inference path: the only communication is the result aggregation below (a barrier, a max over ranks of
the device-measured time, a sum of the unit counts), over ``torch.distributed`` (NCCL on GPUs, gloo on CPU).
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int):
    """Contiguous ``[begin, end)`` share of ``total`` units for ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def aggregate(units: float, seconds: float, device=None):
    """Whole-job ``(total_units, max_seconds)`` over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(units), float(seconds)
    t = torch.tensor([float(units)], dtype=torch.float64, device=device)
    s = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    return float(t.item()), float(s.item())


def exchange(tag: str, payload, timeout_s: float = 900.0):
    """Every rank's ``payload`` (JSON-able) as a list by rank, on every rank, through the rendezvous store: a meeting
    point on the HOST that works whatever state the devices are in (a NCCL collective behind a hung kernel never
    returns).  ``tag`` must be new for every call.  Identity (a one-element list) when not distributed."""
    import datetime
    import json

    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [payload]
    rank, world = dist.get_rank(), dist.get_world_size()
    store = dist.distributed_c10d._get_default_store()
    store.set(f"e55/{tag}/{rank}", json.dumps(payload))
    keys = [f"e55/{tag}/{i}" for i in range(world)]
    store.wait(keys, datetime.timedelta(seconds=timeout_s))
    out = [json.loads(store.get(k)) for k in keys]
    # rank 0 may be the process that hosts the store (env:// without torchrun's agent store): it leaves last
    store.set(f"e55/{tag}/ack/{rank}", "1")
    if rank == 0:
        store.wait([f"e55/{tag}/ack/{i}" for i in range(world)], datetime.timedelta(seconds=timeout_s))
    return out

