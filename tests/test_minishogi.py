"""The native minishogi engine against the reference's own known-answer tests.

The SFEN strings and expected results below are the fixtures of the reference's ``tests/test_checkmate.py``
(9 cases, ``solve_checkmate_dfs(7)``) and ``tests/test_repetition.py`` (8 cases, ``is_repetition()``), which
exercise minishogilib -- the external library this engine stands in for.  The remaining tests are
size-independent rule properties (piece conservation, legality, SFEN round trips) on random playouts.
"""
import random

import numpy as np
import pytest

from erweitern_55_b200 import packing
from erweitern_55_b200.minishogi import Move, Position

START = "rbsgk/4p/5/P4/KGSBR b - 1"

CHECKMATE_KATS = [  # reference tests/test_checkmate.py:7-77
    ("2k2/5/2P2/5/2K2 b G 1", True),
    ("5/5/2k2/5/2K2 b 2GS 1", True),
    ("5/5/2k2/5/2K2 b 2G 1", False),
    ("2k2/5/2B2/5/2K2 b GSBRgsr2p 1", True),
    ("2G1k/5/4G/5/2K2 b P 1", False),
    ("4k/5/4B/5/2K1R b - 1", True),
    ("4k/4p/5/5/K4 b BG 1", True),
    ("5/4k/3pp/5/K4 b RG 1", True),
    ("rbsgk/4p/5/P4/KGSBR b - 1 moves 2e3d 2a2b 4e4d 4a3b 3d4e 5a4a 1e2e 2b2a 5d5c 3a2b 3e3d 3b1d 2e2d 1d3b 4d3c 2b3c 3d3c G*2b 3c2b 2a2b S*4d S*2a G*3e 3b4c 2d2b 2a2b 4d4c 4a4c 4e5d S*3b B*4d R*3a G*4e 4c4d 3e4d 3b2c R*2e B*3d 2e3e 3d4e+ 4d4e 3a3e+ 4e3e R*3a R*4e 2b3c B*1e G*3b 1e3c 3b3c S*4b 3c3b 4b3a 3b3a 3e4d S*3d R*5a 3d4e+ 5d4e R*3e S*3c B*4a 5c5b 3e2e+ 5e5d 2c3b 3c3b 4a3b S*4c 3b4a 4d3d S*4b 4c4b 3a4b 5a4a 4b4a B*4d 1a2a 3d3c R*2d S*2b 2d2b 3c2b 2e2b 4d2b S*4c 5d5e S*3d R*2d 3d4e 2b1a+ 2a3b 1a2a 3b4b 2a4c 4b4c S*4d 4c5b R*5c 5b4b 5e4e 4b3a S*2b 3a4b 2b3c 4b3a 3c2b 3a4b 2b3c 4b3a 3c2b 3a4b 5c4c 4b5b", False),
]

REPETITION_KATS = [  # reference tests/test_repetition.py:7-59
    ("rbsgk/4p/5/P4/KGSBR b - 1", (False, False, False)),
    ("rbsgk/4p/5/P4/KGSBR b - 1 moves 5e4d 1a2b 4d5e 2b1a 5e4d 1a2b 4d5e 2b1a 5e4d 1a2b 4d5e 2b1a", (True, False, False)),
    ("rbsgk/4p/5/P4/KGSBR b - 1 moves 3e2d 3a4b 2e3d 2a2b 4e4d 4a3b 5e4e 5a4a 3d5b 4a5a 5b3d 5a4a 3d5b 4a5a 5b2e 5a4a 2e5b 4a5a 5b3d 5a4a 3d5b", (True, False, False)),
    ("2k2/5/5/5/2K2 b R 1 moves R*3c 3a2a 3c2c 2a3a 2c3c 3a2a 3c2c 2a3a 2c3c 3a2a 3c2c 2a3a 2c3c", (True, False, True)),
    ("rbsgk/4p/5/P4/KGSBR b - 1 moves 4e4d 4a3b 2e3d 3a2b 3e2d 5a4a 5d5c 4a4b 5c5b 4b4d 5e4d G*1d 1e1d 3b1d R*1e 1d3b G*4b R*5d 4d4e 5d3d 4e3d B*3a 4b3b 2a3b 1e1b 1a1b R*1e 1b2a B*1b 2a1a 1b2c 1a2a 2c1b 2a1a 1b2c 1a2a 2c1b 2a1a 1b2c 1a2a 2c1b", (True, False, True)),
    ("3k1/5/2R2/5/2K2 b - 1 moves 3c2c 2a3a 2c3c 3a2a 3c2c 2a3a 2c3c 3a2a 3c2c 2a3a 2c3c 3a2a", (True, True, False)),
    ("rbsgk/4p/5/P4/KGSBR b - 1 moves 5e4d 1a2b 4d5e 2b1a 5e4d 1a2b 4d5e 2b1a", (False, False, False)),
    ("2k2/5/5/5/2K2 b R 1 moves R*3c 3a2a 3c2c 2a3a 2c3c 3a2a 3c2c 2a3a 2c3c 3a2a 3c2c 2a3a", (False, False, False)),
]


@pytest.mark.parametrize("sfen,expected", CHECKMATE_KATS)
def test_checkmate_known_answers(sfen, expected):
    position = Position()
    position.set_sfen(sfen)
    checkmate, move = position.solve_checkmate_dfs(7)
    assert checkmate == expected
    if checkmate:                                     # the returned move is legal and gives check
        assert move in position.generate_moves()


@pytest.mark.parametrize("sfen,expected", REPETITION_KATS)
def test_repetition_known_answers(sfen, expected):
    position = Position()
    position.set_sfen(sfen)
    assert position.is_repetition() == expected


def test_start_position_and_basic_api():
    p = Position()
    p.set_start_position()
    assert p.sfen() == START and p.get_side_to_move() == 0 and p.get_ply() == 0
    moves = p.generate_moves()
    assert len(moves) == 14                            # known perft(1) of minishogi
    assert sorted(m.sfen() for m in moves)[:3] == ["1e1b", "1e1c", "1e1d"]
    q = p.copy(True)
    p.do_move(Move("5d5c"))
    assert p.get_ply() == 1 and p.get_side_to_move() == 1 and q.get_ply() == 0
    assert p.sfen().startswith("rbsgk/4p/P4/5/KGSBR w")
    with pytest.raises(ValueError):
        p.do_move(Move("5d5c"))                        # no longer legal
    with pytest.raises(ValueError):
        p.set_sfen("nonsense")


def _perft(p, depth):
    if depth == 0:
        return 1
    total = 0
    for m in p.generate_moves():
        q = p.copy(True)
        q.do_move(m)
        total += _perft(q, depth - 1)
    return total


def test_perft_start_position():
    """Published minishogi perft counts from the initial position (depth 1-3: 14, 181, 2512)."""
    p = Position()
    assert [_perft(p, d) for d in (1, 2, 3)] == [14, 181, 2512]


def test_drop_rules():
    p = Position()
    p.set_sfen("4k/5/5/P4/K4 b P 1")                   # nifu: no second pawn on file 5
    assert not any(m.sfen().startswith("P*5") for m in p.generate_moves())
    assert not any(m.sfen() in ("P*4a", "P*3a", "P*2a") for m in p.generate_moves())   # no pawn on the last rank
    p.set_sfen("2G1k/5/4G/5/2K2 b P 1")                # reference KAT 5: P*1b would be pawn-drop mate
    assert "P*1b" not in [m.sfen() for m in p.generate_moves()]
    p.set_sfen("4k/4P/5/5/K4 b - 1")                   # a pawn reaching the last rank must promote
    ms = [m.sfen() for m in p.generate_moves()]
    assert "1b1a+" in ms and "1b1a" not in ms


def test_random_playouts_conserve_material_and_round_trip():
    rng = random.Random(5)
    for game in range(20):
        p = Position()
        for _ in range(120):
            rep, my_check, op_check = p.is_repetition()
            moves = p.generate_moves()
            if rep or not moves:
                break
            p.do_move(rng.choice(moves))
            s = p.sfen()
            board, side, hands, _ = s.split()
            counts = {k: 0 for k in "kgsbrp"}
            n = 0                                                   # digits in the hand section are multipliers
            for c in hands:
                if c.isdigit():
                    n = int(c)
                elif c.isalpha():
                    counts[c.lower()] += n or 1
                    n = 0
            for c in board:
                if c.isalpha():
                    counts[c.lower()] += 1
            assert counts == {"k": 2, "g": 2, "s": 2, "b": 2, "r": 2, "p": 2}, s
            q = Position()
            q.set_sfen(s)
            assert q.sfen().split()[:3] == s.split()[:3]
            assert sorted(m.sfen() for m in q.generate_moves()) == sorted(m.sfen() for m in p.generate_moves())


def test_network_input_and_policy_index():
    p = Position()
    p.set_sfen(START + " moves 5d5c 1b1c 5c5b 1c1d")
    bits, scale = p.to_packed_input()
    x = np.asarray(p.to_alphazero_input(), dtype=np.float32).reshape(266, 5, 5)
    assert x.min() >= 0.0 and x.max() <= 1.0                       # network.py:169-170
    b2, s2 = packing.pack_planes(x[None])
    np.testing.assert_array_equal(packing.unpack_planes(b2, s2)[0], x)
    assert x[:10].sum() == 6 and x[10:20].sum() == 6               # six pieces per side on the board
    assert x[33:43].sum() == 6 and x[7 * 33:8 * 33].sum() == 0     # one step back exists, seven steps back does not
    assert np.all(x[264] == 0) and np.allclose(x[265], 4 / 512)    # first player to move, ply 4
    moves = p.generate_moves()
    idx = [p.policy_index(m) for m in moves]
    assert len(set(idx)) == len(idx) and all(0 <= i < 1725 for i in idx)
    # the second player sees the board rotated: mirrored positions give mirrored encodings
    w = Position()
    w.set_sfen(START + " moves 5d5c")
    xw = np.asarray(w.to_alphazero_input(), dtype=np.float32).reshape(266, 5, 5)
    assert np.all(xw[264] == 1)
    assert xw[0].sum() == 1 and xw[0, 4, 0] == 1                   # own king appears bottom-left from its own side


def test_get_inputs_fast_path_equals_the_reference_loop():
    """Network.get_inputs on this package's positions (one native encode) == the reference's per-position loop over
    to_alphazero_input() (network.py:255-274)."""
    from erweitern_55_b200.network import Network
    rng = random.Random(5)
    positions, p = [], Position()
    for _ in range(12):
        moves = p.generate_moves()
        p.do_move(rng.choice(moves))
        positions.append(p.copy(True))
    fast = Network.get_inputs.__get__(type("N", (), {"input_shape": [266, 5, 5]})())(positions)
    slow = np.stack([np.reshape(q.to_alphazero_input(), [266, 5, 5]) for q in positions]).astype(np.float32)
    assert fast.dtype == np.float32 and fast.shape == (12, 266, 5, 5) and np.array_equal(fast, slow)
