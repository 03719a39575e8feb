"""The search kernels of the GPU-resident self-play (csrc/e55_gpu_selfplay.cuh), compiled for the host from the very
source the GPU runs (tests/native/), held to oracle/mcts_oracle.py selection by selection: the CPU-suite twin of
tests/test_gpu_parity.py::test_gpu_search_walks_like_the_oracle, on more and longer searches."""
import os
import shutil
import subprocess

import pytest

from search_walk import HostStepper, walk_like_the_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def stepper(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    lib = str(tmp_path_factory.mktemp("gs") / "libgs_stepper.so")
    cmd = ["g++", "-O2", "-g", "-std=c++20", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-attributes", "-Wno-unknown-pragmas",
           "-Itests/native/shim", "-Ierweitern_55_b200/csrc", "tests/native/gs_stepper.cpp", "-o", lib]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return HostStepper(lib)


REPEAT = ["5e4d", "1a2b", "4d5e", "2b1a", "5e4d", "1a2b", "4d5e", "2b1a", "5e4d", "1a2b"]
# a middle game with pieces in hand on both sides (drops, promotions and captures within reach of the search)
MIDDLE = ["2e1d", "4a3b", "4e4d", "2a2b", "3e3d", "5a5c", "1d3b", "5c2c", "1e1b", "2b1b", "5e4e", "R*1e", "B*2e", "1e1d+"]


@pytest.mark.parametrize("prefix,simulations,searches,full", [
    ([], 800, 3, False),                      # the three walks of the GPU test ...
    (REPEAT, 480, 2, False),
    ([], 640, 2, True),
    ([], 160, 8, False),                      # ... and more: eight moves of one game, the tree reused seven times
    (REPEAT[:6], 320, 4, False),              # repetitions met deeper in the tree, with reuse
    (MIDDLE, 480, 3, False),
    (MIDDLE, 320, 3, True),                   # forced playouts and pruned targets with drops on the board
    (REPEAT, 1024, 1, True),                  # the largest search the kernels accept
])
def test_host_build_of_the_search_walks_like_the_oracle(stepper, prefix, simulations, searches, full):
    selections, terminals, played, ended = walk_like_the_oracle(stepper, prefix, simulations, searches, full)
    assert selections > 0 and (ended or len(played) == len(prefix) + searches)
    if prefix is REPEAT:
        assert terminals > 0                  # the walk did run into fourth occurrences
