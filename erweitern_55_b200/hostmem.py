"""Page-locked result arrays for the host-buffer calls.

``Network.predict`` hands its results back as ordinary ``numpy.ndarray`` objects (network.py:201-216), but for large
batches the memory behind them comes from ``e55_host_alloc`` (cudaHostAlloc): the copy engine writes the 7 MB of
logits of a 1024-position batch straight into the array the caller receives -- no staging copy, no page faults on a
fresh ``np.empty``.  A block returns to a small free list when the last array (or view) on it is garbage-collected
and is handed out again by a later call, so a search loop keeps reusing the same two or three blocks.
"""
from __future__ import annotations

import ctypes as C
import math
import threading

import numpy as np

from . import _capi

_MIN_BYTES = 256 << 10          # smaller results are ordinary numpy arrays
_MAX_CACHED_BYTES = 512 << 20   # free-list cap; beyond it blocks go back to the driver

_F32 = np.dtype(np.float32)
_lock = threading.RLock()      # re-entrant: a garbage collection inside a locked section may run a block's __del__ on the same thread
_free: dict[int, list[int]] = {}
_cached_bytes = 0


def _round(nbytes: int) -> int:
    size = 1 << 18
    while size < nbytes:
        size *= 2
    return size


class _Block:
    """Owner of one pinned allocation; numpy keeps it alive through ``ndarray.base`` (views included)."""

    __slots__ = ("ptr", "size", "__array_interface__")

    def __init__(self, ptr, size, shape, typestr):
        self.ptr, self.size = ptr, size
        self.__array_interface__ = {"data": (ptr, False), "shape": tuple(shape), "typestr": typestr, "version": 3}

    def __del__(self):
        global _cached_bytes
        try:
            with _lock:
                if _cached_bytes + self.size <= _MAX_CACHED_BYTES:
                    _free.setdefault(self.size, []).append(self.ptr)
                    _cached_bytes += self.size
                    return
            _capi.load().e55_host_free(self.ptr)
        except Exception:      # interpreter shutdown
            pass


def empty(shape, dtype=np.float32):
    """Uninitialised array like ``np.empty``; page-locked when it is large enough to matter."""
    global _cached_bytes
    dt = _F32 if dtype is np.float32 else np.dtype(dtype)
    nbytes = int(math.prod(shape)) * dt.itemsize      # (np.prod costs 3 us: a third of a small call's Python time)
    if nbytes < _MIN_BYTES:
        return np.empty(shape, dtype=dt)
    size = _round(nbytes)
    ptr = None
    with _lock:
        lst = _free.get(size)
        if lst:
            ptr = lst.pop()
            _cached_bytes -= size
    if ptr is None:
        out = C.c_void_p()
        if _capi.load().e55_host_alloc(size, C.byref(out)) != _capi.E55_OK or not out.value:
            return np.empty(shape, dtype=dt)      # pinned memory exhausted: pageable still works, just slower
        ptr = out.value
    return np.asarray(_Block(ptr, size, shape, dt.str))


def trim():
    """Give every cached block back to the driver."""
    global _cached_bytes
    with _lock:
        ptrs = [p for lst in _free.values() for p in lst]
        _free.clear()
        _cached_bytes = 0
    for p in ptrs:
        _capi.load().e55_host_free(p)
