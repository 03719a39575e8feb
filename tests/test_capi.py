"""The C-ABI library loads and exports every symbol include/e55.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from erweitern_55_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "e55.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(e55_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert header_symbols() == sorted(_capi.SIGNATURES)


def test_library_exports_every_header_symbol():
    lib = C.CDLL(_capi.lib_path())
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert _capi.load().e55_version() == 1


def test_error_conventions_without_compute():
    lib = _capi.load()
    assert lib.e55_create(None, 0, 7, 128, 266, 16) == _capi.E55_ERR_INVALID
    ctx = C.c_void_p()
    assert lib.e55_create(C.byref(ctx), 0, 7, 100, 266, 16) == _capi.E55_ERR_UNSUPPORTED
    assert b"filters" in lib.e55_last_error(None)
    assert lib.e55_sync(None) == _capi.E55_ERR_INVALID
    assert lib.e55_predict(None, None, 1, None, None) == _capi.E55_ERR_INVALID
    lib.e55_destroy(None)  # no-op


def test_no_cpu_fallback():
    """Without a B200 the product path must fail loudly, never compute on the host."""
    import torch
    from erweitern_55_b200.network import Network
    with pytest.raises(ValueError):
        Network(device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(_capi.E55Error) as ei:
            Network()
        assert ei.value.code == _capi.E55_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "erweitern_55_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include[^\n]*oracle|oracle[./]_ref", src, re.M), f


def test_slice_bookkeeping_of_the_tower_kernel_at_any_launch_length():
    """`e55::tc::SliceBook` (host-callable part of csrc/e55_tc.cuh) against plain counters: the barrier parity the epilogue
    warps wait on before re-using a stem-input buffer stays right past 255, 65535 and 2**20 slices per buffer (a packed
    8-bit version of this bookkeeping once wrapped to the wrong parity and hung launches of more than ~85 rounds per CTA)."""
    lib = _capi.load()
    for slices, seed in ((0, 1), (7, 2), (4 * 256 + 3, 3), (4 * 70000, 4), (5_000_000, 5)):
        bad = C.c_int64(-1)
        assert lib.e55_selftest_slice_book(slices, seed, C.byref(bad)) == 0 and bad.value == 0, (slices, bad.value)
    assert lib.e55_selftest_slice_book(-1, 0, C.byref(bad)) == _capi.E55_ERR_INVALID
