"""The rules oracle (oracle/minishogi_oracle.py, pure Python) pinned on the reference's own fixtures, then the native
engine (csrc/minishogi.hpp behind erweitern_55_b200.minishogi.Position) held against it on random positions: same
legal moves, same repetition verdicts, same mate-search answers."""
import random

import pytest

from erweitern_55_b200.minishogi import Move, Position
from oracle import minishogi_oracle as O
from test_minishogi import CHECKMATE_KATS, REPETITION_KATS, START


@pytest.mark.parametrize("sfen,expected", CHECKMATE_KATS)
def test_oracle_checkmate_known_answers(sfen, expected):          # reference tests/test_checkmate.py:7-77
    pos = O.Position(sfen)
    found, move = pos.solve_checkmate_dfs(7)
    assert found == expected
    if found:
        assert move in pos.generate_moves()


@pytest.mark.parametrize("sfen,expected", REPETITION_KATS)
def test_oracle_repetition_known_answers(sfen, expected):         # reference tests/test_repetition.py:7-59
    assert O.Position(sfen).is_repetition() == expected


def test_oracle_perft_start_position():
    assert [O.perft(O.Position(START), d) for d in (1, 2, 3)] == [14, 181, 2512]


def _playout_positions(seed, games, max_plies):
    """(native position, oracle position) pairs along random games, both fed the same moves."""
    rng = random.Random(seed)
    for _ in range(games):
        native, oracle = Position(), O.Position(START)
        for _ in range(max_plies):
            yield native, oracle
            moves = sorted(m.sfen() for m in native.generate_moves())
            if not moves or native.is_repetition()[0]:
                break
            # captures and drops more often than a uniform choice would: hands fill up, kings get chased
            sharp = [m for m in moves if "*" in m or "+" in m]
            mv = rng.choice(sharp) if sharp and rng.random() < 0.35 else rng.choice(moves)
            native.do_move(Move(mv))
            oracle.do_move(mv)


def test_native_engine_generates_the_oracles_moves():
    n = 0
    for native, oracle in _playout_positions(seed=2024, games=14, max_plies=70):
        assert sorted(m.sfen() for m in native.generate_moves()) == sorted(oracle.generate_moves()), native.sfen()
        assert native.is_repetition() == oracle.is_repetition(), native.sfen(True)
        assert native.sfen().rsplit(" ", 1)[0] == oracle.board_sfen()
        n += 1
    assert n > 300


def test_native_mate_search_agrees_with_the_oracle():
    checked = mates = 0
    for i, (native, oracle) in enumerate(_playout_positions(seed=7, games=30, max_plies=60)):
        if i % 3:
            continue
        found, move = native.solve_checkmate_dfs(3)
        assert found == oracle.solve_checkmate_dfs(3)[0], native.sfen()
        if found:                                                   # the native move is one the oracle accepts as mating
            mates += 1
            assert move.sfen() in oracle.generate_moves()
            undo = oracle._apply(move.sfen())
            assert oracle.in_check(oracle.side) and oracle._defend(2), native.sfen()
            oracle._revert(undo)
        checked += 1
    assert checked > 150 and mates > 0


def test_native_input_planes_and_policy_index_follow_the_documented_layout():
    import numpy as np
    n = 0
    for i, (native, oracle) in enumerate(_playout_positions(seed=99, games=8, max_plies=60)):
        if i % 2:
            continue
        got = np.asarray(native.to_alphazero_input(), dtype=np.float32).reshape(266, 5, 5)
        want = np.asarray(oracle.to_alphazero_input(), dtype=np.float32)
        assert np.array_equal(got, want), (native.sfen(True), np.argwhere(got != want)[:5])
        for m in native.generate_moves():
            assert native.policy_index(m) == oracle.policy_index(m.sfen()), (native.sfen(), m.sfen())
        n += 1
    assert n > 100


@pytest.mark.parametrize("depth,games,every", [(5, 30, 6), (7, 14, 8)])          # 7 = selfplay.py:36's depth
def test_native_deep_mate_search_agrees_with_the_oracle(depth, games, every):
    """Middle-game positions (pieces in hand, exposed kings): mate or no mate within `depth` plies, native against oracle."""
    checked = mates = 0
    rng = random.Random(11 * depth)
    for _ in range(games):
        native, oracle = Position(), O.Position(START)
        for ply in range(70):
            moves = sorted(m.sfen() for m in native.generate_moves())
            if not moves or native.is_repetition()[0]:
                break
            if ply >= 16 and ply % every == 0:
                found, move = native.solve_checkmate_dfs(depth)
                assert found == oracle.solve_checkmate_dfs(depth)[0], (depth, native.sfen())
                if found:
                    assert move.sfen() in moves
                checked += 1
                mates += found
            sharp = [m for m in moves if "*" in m or "+" in m]
            mv = rng.choice(sharp) if sharp and rng.random() < 0.35 else rng.choice(moves)
            native.do_move(Move(mv))
            oracle.do_move(mv)
    assert checked > 50 and mates > 5


def test_native_repetition_verdicts_on_games_steered_into_repetitions():
    """Both sides keep taking their last move back when they can: fourth occurrences and perpetual checks do happen,
    and the native engine's (repetition, my_check, op_check) is the oracle's at every ply."""
    rng = random.Random(5)
    repetitions = perpetual = plies = 0
    starts = [START, START, "2k2/5/5/5/2K2 b R 1", "3k1/5/2R2/5/2K2 b - 1"]      # the last two: a rook chasing a bare king
    for game in range(48):
        sfen = starts[game % 4]
        native, oracle = Position(), O.Position(sfen)
        native.set_sfen(sfen)
        played = []
        for _ in range(60):
            verdict = native.is_repetition()
            assert verdict == oracle.is_repetition(), native.sfen(True)
            plies += 1
            if verdict[0]:
                repetitions += 1
                perpetual += verdict[1] or verdict[2]
                break
            moves = sorted(m.sfen() for m in native.generate_moves())
            if not moves:
                break
            back = played[-2][2:4] + played[-2][0:2] if len(played) >= 2 and "*" not in played[-2] and "+" not in played[-2] else None
            checks = []
            if rng.random() < 0.5:                                   # a checking move, if there is one: perpetual checks
                for m in moves:
                    undo = oracle._apply(m)
                    if oracle.in_check(oracle.side):
                        checks.append(m)
                    oracle._revert(undo)
            mv = back if back in moves and rng.random() < 0.8 else (rng.choice(checks) if checks else rng.choice(moves))
            played.append(mv)
            native.do_move(Move(mv))
            oracle.do_move(mv)
    assert repetitions >= 10 and perpetual >= 1 and plies > 500, (repetitions, perpetual, plies)
