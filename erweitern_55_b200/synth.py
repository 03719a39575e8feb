"""Synthetic network inputs (tier A).

The real plane encoder lives in minishogilib (``Position.to_alphazero_input``, called at reference
``erweitern_55/network.py:258``) which is not available here; the reference only pins the shape
``[N,266,5,5]``, dtype float32 and range [0,1] (``network.py:85,169-170``).  Tier-A inputs follow that
contract: sparse 0/1 history planes plus two constant planes (side to move, ply/512).
"""
from __future__ import annotations

import numpy as np

from .weights import BOARD, IN_PLANES


def synthetic_planes(n: int, seed: int = 1000, in_planes: int = IN_PLANES, density: float = 0.1):
    """float32 ``[n, in_planes, 5, 5]``: Bernoulli(density) history planes, two constant planes."""
    rng = np.random.default_rng(seed)
    x = (rng.random((n, in_planes, BOARD, BOARD)) < density).astype(np.float32)
    x[:, 0] = rng.integers(0, 2, size=(n, 1, 1)).astype(np.float32)
    x[:, 1] = (rng.integers(0, 512, size=(n, 1, 1)) / 512.0).astype(np.float32)
    return x


def synthetic_legal_mask(n: int, seed: int = 2000, policy_size: int = 1725, moves: int = 30):
    """uint8 ``[n, policy_size]`` with ``moves`` random legal entries per row (at least one)."""
    rng = np.random.default_rng(seed)
    mask = np.zeros((n, policy_size), dtype=np.uint8)
    for i in range(n):
        mask[i, rng.choice(policy_size, size=moves, replace=False)] = 1
    return mask
