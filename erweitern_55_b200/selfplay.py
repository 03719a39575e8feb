"""Self-play games with the reference's surface: ``run`` and ``random_play`` of ``erweitern_55/selfplay.py:10-169``.

``run(nn, search, ...)`` plays ONE game from the start position and returns a ``GameRecord`` like the reference:
every ply first checks repetition / stalemate (selfplay.py:17-31), then tries the 7-ply mate search
(selfplay.py:35-42), otherwise searches (with the playout-cap oscillation of selfplay.py:45-61 when enabled) and
samples from the visit counts for the first ``num_sampling_moves`` plies.  ``run_many`` is this package's addition:
many games at once on the native driver (``e55_selfplay_run_ex``), one host thread per game, leaf batches of all games
coalesced per GPU by the evaluation service -- the way a self-play shard should be run on a B200.
"""
from __future__ import annotations

import ctypes as C
import random
import time
from datetime import datetime

import numpy as np

from . import _capi, gamerecord, minishogi


def _game_over(position, record):
    """Termination tests of selfplay.py:17-31; sets ``record.winner`` when the game has ended."""
    is_repetition, my_check_repetition, op_check_repetition = position.is_repetition()
    side = position.get_side_to_move()
    if my_check_repetition:
        record.winner = 1 - side
    elif op_check_repetition:
        record.winner = side
    elif is_repetition:
        record.winner = 1
    elif len(position.generate_moves()) == 0:
        record.winner = 1 - side
    else:
        return False
    return True


def _trim_mate(record):
    record.ply -= 2
    for name in ("sfen_kif", "mcts_result", "learning_target_plys"):
        setattr(record, name, getattr(record, name)[:-2])


def run(nn, search, verbose=False, num_sampling_moves=10, max_moves=512,
        playout_cap_oscillation={'enable': False, 'N': 800, 'n': 128, 'frac': 0.25},
        search_checkmate=True, stop_with_checkmate=False, trim_checkmate=False):
    position = minishogi.Position()
    position.set_start_position()
    record = gamerecord.GameRecord()
    cap = playout_cap_oscillation

    for _ in range(max_moves):
        if _game_over(position, record):
            break
        started = time.time()
        checkmate, checkmate_move = position.solve_checkmate_dfs(7) if search_checkmate else (None, None)
        root = None
        if checkmate:
            next_move = checkmate_move
            if not stop_with_checkmate:
                search.mcts.clear()
        else:
            if cap['enable']:
                full = np.random.rand() < cap['frac']          # a full search produces a learning target
                cfg = search.config
                cfg.simulation_num = cap['N'] if full else cap['n']
                cfg.forced_playouts = cfg.use_dirichlet = cfg.target_pruning = full
                cfg.reuse_tree = cfg.immediate = not full
            root = search.run(position, nn)
            if position.get_ply() < num_sampling_moves:
                next_move = search.softmax_sample(root, 1.0)
            else:
                next_move = search.best_move(root)
        if verbose:
            if checkmate:
                print('checkmate!')
            else:
                search.print(root)
        elapsed = time.time() - started

        position.do_move(next_move)
        record.sfen_kif.append(next_move.sfen())
        if checkmate:
            record.mcts_result.append((1, 1.0, [(checkmate_move.sfen(), 1)]))
            record.learning_target_plys.append(record.ply)
        else:
            record.mcts_result.append(search.dump(root))
            if not cap['enable'] or search.config.simulation_num >= cap['N']:
                record.learning_target_plys.append(record.ply)
        record.ply += 1

        if verbose:
            print('--------------------')
            position.print()
            print(next_move)
            print('time:', elapsed)
            print('--------------------')
        if checkmate and stop_with_checkmate:
            record.winner = 1 - position.get_side_to_move()
            if trim_checkmate:
                _trim_mate(record)
            break

    record.timestamp = int(datetime.now().timestamp())
    return record


def random_play(max_moves=512, stop_with_checkmate=False, trim_checkmate=False):
    position = minishogi.Position()
    position.set_start_position()
    record = gamerecord.GameRecord()

    for _ in range(max_moves):
        if _game_over(position, record):
            break
        checkmate, checkmate_move = position.solve_checkmate_dfs(7)
        next_move = checkmate_move if checkmate else random.choice(position.generate_moves())
        position.do_move(next_move)
        record.sfen_kif.append(next_move.sfen())
        record.learning_target_plys.append(record.ply)
        record.mcts_result.append((1, 1.0 if checkmate else 0.5, [(next_move.sfen(), 1)]))
        record.ply += 1
        if checkmate and stop_with_checkmate:
            record.winner = 1 - position.get_side_to_move()
            if trim_checkmate:
                _trim_mate(record)
            break

    record.timestamp = int(datetime.now().timestamp())
    return record


def run_many(nn, threads=0, total_moves=1000, simulations=800, leaf_batch=16, mate_depth=7, reuse_tree=True,
             use_dirichlet=True, num_sampling_moves=10, max_moves=512, seed=1, records=True, games_per_thread=0,
             playout_cap_oscillation=None):
    """Play games on ``threads`` host threads (0: cores - 2) working through a pool of ``threads x games_per_thread``
    games (0: 12 per thread; a service batch of about 74 positions per thread suits that) kept in flight on the asynchronous evaluation service, until ``total_moves`` plies were played; the service of ``nn`` must be running
    (``nn.start_service``).  Returns ``(stats dict, [GameRecord of each finished game])``."""
    lib = _capi.load()
    cfg = _capi.SelfplayConfig()
    lib.e55_selfplay_default_config(C.byref(cfg))
    cfg.threads, cfg.total_moves, cfg.simulations, cfg.leaf_batch = threads, total_moves, simulations, leaf_batch
    cfg.games_per_thread = games_per_thread
    cfg.mate_depth, cfg.reuse_tree, cfg.use_dirichlet = mate_depth, int(reuse_tree), int(use_dirichlet)
    cfg.sampling_moves, cfg.max_moves, cfg.seed = num_sampling_moves, max_moves, seed
    if playout_cap_oscillation and playout_cap_oscillation.get("enable"):     # selfplay.py:45-61
        cap = playout_cap_oscillation
        cfg.cap_enable, cfg.cap_full, cfg.cap_fast, cfg.cap_frac = 1, int(cap["N"]), int(cap["n"]), float(cap["frac"])
    games = []
    if records:
        def sink(_user, text):
            games.append(gamerecord.GameRecord.from_native(text.decode(), int(datetime.now().timestamp())))
        cfg.sink = _capi.GAME_SINK(sink)
    st = _capi.SelfplayStats()
    _capi.check(nn._ctx, lib.e55_selfplay_run_ex(nn._ctx, C.byref(cfg), C.byref(st)))
    stats = {"seconds": st.seconds, "moves": st.moves, "evals": st.evals, "games": st.games,
             "mean_game_plies": st.mean_game_plies, "mean_gpu_batch": st.mean_gpu_batch,
             "moves_per_s": st.moves / st.seconds, "evals_per_s": st.evals / st.seconds}
    return stats, games
