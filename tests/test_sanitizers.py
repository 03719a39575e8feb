"""Memory-error and data-race checks of the host-side native code (no GPU): the minishogi engine, the search tree, the
C API the Python classes bind and the pooled self-play driver are built with AddressSanitizer + UBSan and with
ThreadSanitizer against the stub network of tools/host_prof.cpp and driven the way mcts.py / selfplay.py drive them.
compute-sanitizer is closed on the GPU pool (profiles/r02_compute_sanitizer_closed.log); of the CUDA side, the search
kernels of the GPU-resident self-play are thread-per-game code without tensor-core or shared-memory machinery, so the
very same source is compiled for the host (tests/native/) and run under AddressSanitizer + UBSan here, its trees
audited after every round and its games replayed on the rules oracle."""
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


# ---- the search kernels of the GPU-resident self-play (csrc/e55_gpu_selfplay.cuh), host build ------------------------
GS_SOURCES = ["-ffp-contract=off", "-Wno-attributes", "-Itests/native/shim", "-Ierweitern_55_b200/csrc", "tests/native/gs_host.cpp"]


def _gs_run(exe, *, games, moves, simulations, mate_depth=7, reuse_tree=1, noise=1, max_moves=512, seed=1, sharp=6.0,
            cap=(0, 800, 128, 0.25), record_cap=64, records_out="-"):
    import json
    out = _run(exe, *map(str, (games, moves, simulations, mate_depth, reuse_tree, noise, max_moves, seed, sharp, *cap, record_cap, records_out)))
    return json.loads(out.strip().splitlines()[-1])


def _gs_clean(st):
    assert st["audit_failures"] == 0, st["first_audit_failure"]
    assert st["bad_games"] == 0, st["first_bad_game"]
    assert st["pool_overflows"] == 0 and st["records_checked"] > 0
    assert st["evals"] == st["evals_by_slots"] and st["audit_rounds"] == st["rounds"]


def test_gpu_search_kernels_on_the_host_under_asan_ubsan(tmp_path):
    """init_kernel / select_kernel / expand_kernel as the GPU runs them (same source, tests/native/shim), a pseudo-network
    in the tower kernel's place: memory-clean, every tree consistent after every round (no virtual loss left behind, node
    counts = parent edge counts, links and pools in range), every finished game legal on the host engine AND, decoded by the
    product's record decoder, on the independent rules oracle; mate-search moves are mates."""
    import numpy as np
    from erweitern_55_b200.network import decode_gpu_records
    from oracle.minishogi_oracle import Position
    exe = _build(tmp_path, "gs_host_san", ["-fsanitize=address,undefined", "-fno-sanitize-recover=all"], GS_SOURCES)
    rec_file = str(tmp_path / "records.bin")
    st = _gs_run(exe, games=24, moves=1500, simulations=160, seed=2, records_out=rec_file)
    _gs_clean(st)
    assert st["moves"] >= 1500 and st["games"] >= 20 and st["mate_moves"] > 0 and st["kept_nodes"] > 0
    assert st["searches_suspended"] == 0 and st["records_truncated"] == 0 and st["ended_without_moves"] > 0
    buf = np.fromfile(rec_file, dtype=np.uint16).reshape(-1, 16384)
    games = decode_gpu_records(buf, len(buf))
    assert len(games) == st["records_checked"]
    from erweitern_55_b200 import minishogi as native
    mates = 0
    for rec in games[:12]:                                             # the independent (pure-Python) rules on a dozen games
        p, pn = Position(), native.Position()
        assert rec.complete and rec.ply == len(rec.sfen_kif) == len(rec.mcts_result)
        for m, (sum_n, q, visits) in zip(rec.sfen_kif, rec.mcts_result):
            legal = set(p.generate_moves())
            assert m in legal and {v for v, _ in visits} <= legal and 0.0 <= q <= 1.0, (m, p.board_sfen())
            assert sum_n == sum(n for _, n in visits) and dict(visits).get(m, 0) > 0
            if (sum_n, visits) == (1, [(m, 1)]):                       # a mate-search move (selfplay.py:82-84)
                mates += 1
                assert pn.solve_checkmate_dfs(7)[0]
            else:
                assert sum_n >= 160
            p.do_move(m)
            pn.do_move(native.Move(m))
        rep, my_check, op_check = p.is_repetition()
        if rep:
            assert rec.winner == (1 - p.side if my_check else p.side if op_check else 1)
        elif not p.generate_moves():
            assert rec.winner == 1 - p.side
        else:
            assert rec.ply == 512 and rec.winner == 2
    assert mates > 0
    # sharper priors (deep, narrow trees), full-size searches, no mate search, no tree reuse, length limit, playout-cap oscillation
    _gs_clean(_gs_run(exe, games=8, moves=300, simulations=800, seed=3, sharp=20.0))
    st = _gs_run(exe, games=16, moves=1200, simulations=160, mate_depth=5, noise=0, max_moves=40, seed=4, sharp=2.0, cap=(1, 160, 32, 0.25),
                 records_out=rec_file)
    _gs_clean(st)
    assert st["ended_by_length"] > 0
    # playout-cap oscillation in the records (selfplay.py:45-61,90-91): only the fully searched plies (about a quarter) and
    # the mate-search moves are learning targets; a fast search records what its 32 simulations (or the reused tree) saw
    buf = np.fromfile(rec_file, dtype=np.uint16).reshape(-1, 16384)
    games = decode_gpu_records(buf, len(buf))
    plies = sum(g.ply for g in games)
    targets = sum(len(g.learning_target_plys) for g in games)
    assert 0.1 * plies < targets < 0.6 * plies, (targets, plies)
    for g in games:
        assert g.complete and all(0 <= t < g.ply for t in g.learning_target_plys)
        for t, (sum_n, _, visits) in enumerate(g.mcts_result):
            if t in g.learning_target_plys and visits != [(g.sfen_kif[t], 1)]:
                assert sum_n >= 100, (t, sum_n)                          # a full search of 160 simulations, pruned (Wu 2019)
    _gs_clean(_gs_run(exe, games=16, moves=800, simulations=160, mate_depth=0, reuse_tree=0, seed=5, sharp=0.0))


@pytest.mark.parametrize("nodes,edges,words", [(1024, 16384, 512), (256, 65536, 16384)])
def test_gpu_search_kernels_with_small_pools(tmp_path, nodes, edges, words):
    """The same with tree pools and record buffers shrunk at compile time: searches are suspended at 90 % of a pool all the
    time (mcts.py:71-73) and records are truncated all the time -- and still no descent is ever abandoned, no tree is left
    inconsistent, no record is malformed (truncated ones carry their flag), nothing is read or written out of bounds."""
    exe = _build(tmp_path, "gs_host_small", ["-fsanitize=address,undefined", "-fno-sanitize-recover=all", f"-DE55_GS_MAX_NODES={nodes}",
                                             f"-DE55_GS_MAX_EDGES={edges}", f"-DE55_GS_RECORD_WORDS={words}"], GS_SOURCES)
    st = _gs_run(exe, games=16, moves=1200, simulations=800, seed=11, sharp=8.0)
    _gs_clean(st)
    assert st["searches_suspended"] > 100 and st["max_nodes"] <= nodes and st["max_edges"] <= edges
    assert (st["records_truncated"] > 0) == (words < 16384) and st["truncated_records_seen"] <= st["records_truncated"]
    st = _gs_run(exe, games=16, moves=1200, simulations=160, seed=12, sharp=100.0, cap=(1, 800, 160, 0.3))
    _gs_clean(st)
    assert st["searches_suspended"] > 0
