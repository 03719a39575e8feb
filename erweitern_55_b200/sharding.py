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
