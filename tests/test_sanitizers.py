"""Memory-error and data-race checks of the host-side native code (no GPU): the minishogi engine, the search tree, the
C API the Python classes bind and the pooled self-play driver are built with AddressSanitizer + UBSan and with
ThreadSanitizer against the stub network of tools/host_prof.cpp and driven the way mcts.py / selfplay.py drive them.
(The CUDA side cannot be checked this way; compute-sanitizer is not available on the GPU pool.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["tools/host_prof.cpp", "erweitern_55_b200/csrc/e55_game.cpp"]
BASE = ["g++", "-O1", "-g", "-std=c++20", "-march=x86-64-v3", "-pthread", "-fno-omit-frame-pointer", "-I."]


PACKER_SOURCES = ["-Ierweitern_55_b200/csrc", "tools/pack_check.cpp", "erweitern_55_b200/csrc/e55_host.cpp"]


def _build(tmp_path, name, flags, sources=SOURCES):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / name)
    r = subprocess.run(BASE + flags + sources + ["-o", exe], cwd=ROOT, capture_output=True, text=True, timeout=300)
    if r.returncode != 0 and ("cannot find" in r.stderr or "unrecognized" in r.stderr):
        pytest.skip("sanitizer runtime not installed: " + r.stderr.strip().splitlines()[-1])
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _run(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
    report = r.stdout[-1500:] + r.stderr[-3000:]
    if "unexpected memory mapping" in r.stderr or "Shadow memory range interleaves" in r.stderr or "failed to allocate" in r.stderr:
        pytest.skip("the sanitizer runtime cannot start in this environment: " + r.stderr.strip().splitlines()[0])
    assert r.returncode == 0 and "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, report
    return r.stdout


def test_engine_and_tree_api_under_asan_ubsan(tmp_path):
    exe = _build(tmp_path, "host_prof_asan", ["-fsanitize=address,undefined"])
    out = _run(exe, "fuzz", "12", "48")                       # positions, SFEN, mate search, immediate-mode tree with reuse
    assert "ok" in out
    out = _run(exe, "3", "400", "7", "64", "3")                # pooled self-play driver: 3 threads x 3 games, 400 moves
    assert "rc 0: 400 moves" in out


def test_selfplay_driver_under_tsan(tmp_path):
    exe = _build(tmp_path, "host_prof_tsan", ["-fsanitize=thread"])
    out = _run(exe, "4", "300", "7", "48", "3")                # 4 threads stealing games from each other
    assert "rc 0: 300 moves" in out


@pytest.mark.parametrize("sanitizer", ["address,undefined", "thread"])
def test_host_packer_under_sanitizers(tmp_path, sanitizer):
    """The AVX2 plane packer and its thread pool (csrc/e55_host.cpp) against a scalar restatement: ragged batch, constant
    planes, negative zeros, planes that cannot be packed, main thread lending a hand."""
    exe = _build(tmp_path, "pack_check", ["-fsanitize=" + sanitizer], PACKER_SOURCES)
    assert "0 mismatches: ok" in _run(exe, "5", "9")
