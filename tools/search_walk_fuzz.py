"""Long differential run of the GPU search kernels (host build, tests/native/gs_stepper.cpp) against oracle/mcts_oracle.py:
random game prefixes, search sizes, tree reuse and playout-cap "full" searches, selection by selection
(tests/search_walk.py).  No GPU needed.   python tools/search_walk_fuzz.py <seconds> [first seed]"""
import os
import random
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from search_walk import HostStepper, walk_like_the_oracle  # noqa: E402
from oracle.minishogi_oracle import Position  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lib = f"/tmp/libgs_stepper_fuzz_{os.getpid()}.so"
    subprocess.run(["g++", "-O2", "-std=c++20", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-attributes", "-Wno-unknown-pragmas",
                    "-Itests/native/shim", "-Ierweitern_55_b200/csrc", "tests/native/gs_stepper.cpp", "-o", lib], cwd=ROOT, check=True)
    stepper = HostStepper(lib)
    t0, walks, selections, terminals = time.time(), 0, 0, 0
    seed = seed0
    while time.time() - t0 < seconds:
        seed += 1
        rng = random.Random(seed)
        pos, prefix = Position(), []
        for _ in range(rng.choice([0, 2, 6, 10, 16, 24, 34])):
            moves = sorted(pos.generate_moves())
            if not moves or pos.is_repetition()[0]:
                break
            m = rng.choice(moves)
            pos.do_move(m)
            prefix.append(m)
        if not pos.generate_moves() or pos.is_repetition()[0]:
            prefix = prefix[:-1]                                      # the game must not be over before the search
        heavy = os.environ.get("WALK_FUZZ_HEAVY") == "1"             # larger searches, mostly "full" ones (forced playouts, pruning)
        sims = rng.choice([320, 480, 800, 1024] if heavy else [16, 32, 64, 160, 320])
        searches = rng.choice([1, 2] if heavy else [1, 2, 3, 5])
        full = rng.random() < (0.7 if heavy else 0.3)
        try:
            s, t, played, ended = walk_like_the_oracle(stepper, prefix, sims, searches, full, expect_pruning=False)
        except AssertionError:
            print("MISMATCH", dict(seed=seed, prefix=prefix, simulations=sims, searches=searches, full=full), flush=True)
            raise
        walks += 1; selections += s; terminals += t
        if walks % 20 == 0:
            print(f"{walks} walks, {selections} selections on the oracle's path, {terminals} terminal leaves, {time.time() - t0:.0f} s", flush=True)
    print(f"done: {walks} walks, {selections} selections, {terminals} terminal leaves, seeds {seed0 + 1}..{seed}: no mismatch", flush=True)


if __name__ == "__main__":
    main()
