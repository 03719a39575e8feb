"""Bit-packed network input.

The reference feeds float32 planes ``[N,266,5,5]`` (``network.py:255-274``): 26 600 bytes per position, almost
all of them 0.0 or 1.0.  The packed form keeps, per plane, one occupancy bit per square and the single value the
occupied squares hold: ``bits uint32 [N,C]`` (bit ``y*5+x``) and ``scale float32 [N,C]`` = 2 128 bytes per
position.  It is exact for every plane whose non-zero squares share one value (occupancy planes, and the
constant count / side-to-move / ply planes of the AlphaZero encoding); anything else is rejected.
"""
from __future__ import annotations

import numpy as np

_WEIGHTS = (1 << np.arange(25, dtype=np.uint32)).astype(np.uint32)


def pack_planes(images):
    """float planes ``[N,C,5,5]`` -> ``(bits uint32 [N,C], scale float32 [N,C])``; raises if not representable."""
    x = np.ascontiguousarray(np.asarray(images, dtype=np.float32))
    if x.ndim != 4 or x.shape[2:] != (5, 5):
        raise ValueError(f"expected [N,C,5,5] planes, got {list(x.shape)}")
    flat = x.reshape(x.shape[0], x.shape[1], 25)
    nz = flat != 0
    scale = np.abs(flat).max(axis=2) * np.where((flat.min(axis=2) < 0), -1.0, 1.0).astype(np.float32)
    scale = scale.astype(np.float32)
    if not np.array_equal(np.where(nz, scale[:, :, None], np.float32(0)), flat):
        raise ValueError("a plane holds more than one distinct non-zero value and cannot be bit-packed")
    bits = (nz.astype(np.uint32) * _WEIGHTS).sum(axis=2, dtype=np.uint32)
    return np.ascontiguousarray(bits), np.ascontiguousarray(scale)


def unpack_planes(bits, scale):
    """Inverse of :func:`pack_planes` (host reference of what the kernel expands)."""
    bits = np.asarray(bits, dtype=np.uint32)
    scale = np.asarray(scale, dtype=np.float32)
    occ = (bits[:, :, None] >> np.arange(25, dtype=np.uint32)) & 1
    return np.where(occ.astype(bool), scale[:, :, None], np.float32(0)).reshape(bits.shape[0], bits.shape[1], 5, 5)
