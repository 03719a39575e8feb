"""The search stack above the network -- tree (minishogilib.MCTS counterpart), mcts.MCTS.run, selfplay.run,
the USI front-end -- on CPU with a stand-in network object (tree operations are host code; the GPU runs of the
same flows are in test_gpu_parity.py).  Where /root/reference is present, the reference's own unmodified
mcts.py + selfplay.py are run on the import shims of erweitern_55_b200/compat."""
import io
import os
import sys

import numpy as np
import pytest

from erweitern_55_b200 import gamerecord, mcts, selfplay, usi
from erweitern_55_b200.minishogi import MCTS as Tree, Move, Position

START = "rbsgk/4p/5/P4/KGSBR b - 1"


class FakeNetwork:
    """The reference's Network surface (network.py:201-216,255-274) with a deterministic pseudo-network behind it."""

    input_shape = [266, 5, 5]

    def __init__(self, value=None, flat_policy=False):
        self.calls = []
        self._value = value
        self._flat = flat_policy

    def get_input(self, position, dim_4=False):
        shape = [1] + self.input_shape if dim_4 else self.input_shape
        return np.reshape(position.to_alphazero_input(), shape)

    def get_inputs(self, positions):
        ins = np.zeros([len(positions)] + self.input_shape, dtype="float32")
        for i, position in enumerate(positions):
            ins[i] = self.get_input(position)
        return ins

    def predict(self, images):
        x = np.asarray(images, dtype=np.float32)
        self.calls.append(x.shape[0])
        key = x.reshape(x.shape[0], -1) @ np.cos(np.arange(6650, dtype=np.float32))
        policy = np.sin(np.outer(1.0 + key, np.arange(1725, dtype=np.float32) * 0.37)).astype(np.float32)
        value = np.tanh(0.3 * np.sin(key))[:, None].astype(np.float32)
        if self._value is not None:
            value[:] = self._value
        if self._flat:
            policy[:] = 0.0
        return policy, value


def _config(simulations=64, **kw):
    cfg = mcts.Config()
    cfg.simulation_num = simulations
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def test_tree_primitives_keep_their_books():
    p = Position()
    tree = Tree(0.01)
    root = tree.set_root(p, True)
    assert root == 0 and not tree.expanded(root) and tree.best_move(root) is None
    rng = np.random.default_rng(0)
    tree.evaluate(root, p, rng.standard_normal(1725).astype(np.float32), 0.5)
    assert tree.expanded(root) and tree.get_playouts(root, True) == 0
    for _ in range(20):
        leaves, poss = [], []
        for _b in range(8):
            q = p.copy(True)
            leaves.append(tree.select_leaf(root, q, False))
            poss.append(q)
            assert q.get_ply() >= 1                                      # the leaf position was advanced
        for leaf, q in zip(leaves, poss):
            tree.evaluate(leaf, q, rng.standard_normal(1725).astype(np.float32), float(rng.random()))
        for leaf in leaves:
            tree.backpropagate(leaf)
    assert tree.get_playouts(root, True) == 160 and tree.get_playouts(root, False) == 160
    sum_n, q, visits = tree.dump(root, False, True)
    assert sum_n == 160 == sum(n for _, n in visits) and 0.0 <= q <= 1.0
    legal = {m.sfen() for m in p.generate_moves()}
    assert {m for m, _ in visits} <= legal and all(n > 0 for _, n in visits)
    assert len(tree.dump(root, False, False)[2]) == len(legal)           # remove_zeros=False lists every move
    best = tree.best_move(root)
    assert best.sfen() == max(visits, key=lambda t: t[1])[0]
    pv, win = tree.info(root)
    assert pv[0] == best and 0.0 <= win <= 1.0
    assert tree.softmax_sample(root, 1.0).sfen() in legal
    assert tree.softmax_sample_among_top_moves(root, 0.01, 10.0).sfen() in legal
    assert 0 < tree.get_usage() < 1 and tree.get_nodes() > 100
    assert tree.visualize(root, 5).startswith("digraph") and "N" in tree.describe(root)
    # tree reuse: after playing the best move its subtree (and nothing else) becomes the tree
    child_visits = dict(visits)[best.sfen()]
    nodes_before = tree.get_nodes()
    p.do_move(best)
    root = tree.set_root(p, True)
    assert tree.expanded(root) and tree.get_playouts(root, False) == child_visits and tree.get_nodes() < nodes_before
    root = tree.set_root(p, False)                                       # reuse_tree off: empty tree
    assert not tree.expanded(root) and tree.get_nodes() == 1
    q = Position()
    q.set_sfen("2k2/5/5/5/2K2 b R 1")
    tree.set_root(p, True); tree.evaluate(0, p, np.zeros(1725, np.float32), 0.5)
    assert not tree.expanded(tree.set_root(q, True))                     # another game: nothing is kept


def test_values_flip_sign_per_ply_and_virtual_loss_spreads_a_batch():
    p = Position()
    tree = Tree(0.01)
    root = tree.set_root(p, False)
    tree.evaluate(root, p, np.zeros(1725, np.float32), 0.5)
    poss = [p.copy(True) for _ in range(14)]
    leaves = [tree.select_leaf(root, q, False) for q in poss]
    assert len(set(leaves)) == 14                                        # uniform priors + virtual loss: 14 different children
    for leaf, q in zip(leaves, poss):
        tree.evaluate(leaf, q, np.zeros(1725, np.float32), 0.9)          # good for the side to move at the leaf
    for leaf in leaves:
        tree.backpropagate(leaf)
    sum_n, q, visits = tree.dump(root, False, True)
    assert sum_n == 14 and abs(q - 0.1) < 1e-6                           # ... so bad for the root's side


def test_search_finds_the_mate_in_one():
    p = Position()
    p.set_sfen("2k2/5/2P2/5/2K2 b G 1")                                   # reference tests/test_checkmate.py:8
    found, mate = p.solve_checkmate_dfs(7)
    assert found
    search = mcts.MCTS(_config(512, reuse_tree=False))
    root = search.run(p, FakeNetwork(value=0.0, flat_policy=True))    # knows nothing: the tree has to find it
    assert search.best_move(root) == mate
    pv, win = search.mcts.info(root)
    assert win > 0.9


def test_mcts_run_contract():
    nn = FakeNetwork()
    p = Position()
    search = mcts.MCTS(_config(64))
    root = search.run(p, nn)
    assert nn.calls == [1] + [16] * 4                                     # root batch 1, then leaf batches of 16 (mcts.py:58,102)
    assert search.mcts.get_playouts(root, True) == 64
    sum_n, q, visits = search.dump(root)
    assert sum_n == 64
    # tree reuse carries the subtree to the next move, and immediate mode returns once the root has its visits
    best = search.best_move(root)
    carried = dict(visits)[best.sfen()]
    p.do_move(best)
    search.run(p, nn, set_root_only=True)
    search.config.immediate = True
    search.config.simulation_num = max(search.mcts.get_playouts(0, True), 1)   # visits below the carried node
    nn.calls.clear()
    root = search.run(p, nn)
    assert nn.calls == [] and search.mcts.get_playouts(root, False) == carried
    # set_root_only (usi.py:146-148 when pondering is off) moves the root without searching
    search.config.immediate = False
    assert search.run(p, nn, set_root_only=True) is None and nn.calls == []
    # stop() ends a running search early; a time limit does too
    search.config.simulation_num = 10 ** 9
    search.config.reuse_tree = False
    import threading
    import time
    t = threading.Thread(target=search.run, args=(p, nn))
    t.start()
    time.sleep(0.3)
    search.stop()
    t.join(timeout=10)
    assert not t.is_alive()
    t0 = time.time()
    search.run(p, nn, timelimit=200)
    assert time.time() - t0 < 5
    # forced playouts + target pruning (selfplay.py:47-52) still account for every simulation or prune some
    cfg = _config(128, forced_playouts=True, use_dirichlet=True, target_pruning=True, reuse_tree=False)
    search = mcts.MCTS(cfg)
    root = search.run(Position(), nn)
    pruned = search.dump(root)
    cfg.target_pruning = False
    full = search.dump(root)
    assert full[0] == 128 and pruned[0] <= 128 and len(pruned[2]) <= len(full[2])


def _check_record(rec):
    assert rec.ply == len(rec.sfen_kif) == len(rec.mcts_result) and rec.winner in (0, 1, 2)
    p = Position()
    for move, (sum_n, q, visits) in zip(rec.sfen_kif, rec.mcts_result):
        legal = {m.sfen() for m in p.generate_moves()}
        assert move in legal and {m for m, _ in visits} <= legal
        assert sum_n == sum(n for _, n in visits) and 0.0 <= q <= 1.0
        p.do_move(Move(move))
    return p


def test_selfplay_run_plays_a_legal_game_and_records_it():
    np.random.seed(1)
    rec = selfplay.run(FakeNetwork(), mcts.MCTS(_config(32)), max_moves=40)
    end = _check_record(rec)
    assert rec.ply > 4 and rec.learning_target_plys == list(range(rec.ply)) and rec.timestamp > 0
    if rec.winner != 2:                                                   # decided games end in a terminal position
        assert not end.generate_moves() or end.is_repetition()[0]
    capped = selfplay.run(FakeNetwork(), mcts.MCTS(_config(32)), max_moves=12, search_checkmate=False,
                          playout_cap_oscillation={'enable': True, 'N': 48, 'n': 16, 'frac': 0.5})
    _check_record(capped)
    assert set(capped.learning_target_plys) < set(range(capped.ply))     # only the full searches are targets
    assert isinstance(rec.to_dict()["mcts_result"], list)


def test_random_play_and_native_record_text():
    import random
    random.seed(3)
    rec = selfplay.random_play(max_moves=60)
    _check_record(rec)
    text = "winner 1 ply 2\n5d5c 32 0.512 2 5d5c 30 5e4d 2\n1b1c 1 1.0 1 1b1c 1\n"
    got = gamerecord.GameRecord.from_native(text, 7)
    assert (got.winner, got.ply, got.sfen_kif, got.timestamp) == (1, 2, ["5d5c", "1b1c"], 7)
    assert got.mcts_result[0] == (32, 0.512, [("5d5c", 30), ("5e4d", 2)])


def test_usi_session():
    out = io.StringIO()
    engine = usi.USI(nn=FakeNetwork(), stdout=out)
    script = ["usi", "isready", "setoption name ponder value false", "setoption name softmax_sampling_moves value 2",
              "usinewgame", "position startpos", "go btime 3000 wtime 3000 byoyomi 0",
              "position sfen 2k2/5/2P2/5/2K2 b G 1", "go btime 1000 wtime 1000 byoyomi 100",
              "position sfen 4k/5/5/5/K4 w - 1 moves 1a1b", "d", "bogus", "quit"]
    engine._in = io.StringIO("\n".join(script) + "\n")
    engine.start()
    lines = out.getvalue().splitlines()
    assert lines[:4] == ["id name erweitern_55", "id author nyashiki", "usiok", "readyok"]
    assert engine.option["ponder"] is False and engine.option["softmax_sampling_moves"] == 2
    best = [l.split()[1] for l in lines if l.startswith("bestmove")]
    assert len(best) == 2
    assert best[0] in {m.sfen() for m in Position().generate_moves()}
    assert "info string think time 300" in lines and any(l.startswith("info depth") for l in lines)
    assert best[1] == "G*3b"                                              # the mate search answers without thinking
    assert any(l.startswith("ERROR: Unknown command.") for l in lines)
    assert engine.ponder_thread is None


REFERENCE = "/root/reference/erweitern_55"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only mounted in the build container")
def test_reference_mcts_and_selfplay_run_unmodified_on_the_shims():
    from erweitern_55_b200 import compat
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for name in ("mcts", "selfplay", "network", "gamerecord", "minishogilib"):
        sys.modules.pop(name, None)
    sys.path[:0] = [compat.PATH, REFERENCE]
    try:
        import mcts as ref_mcts
        import selfplay as ref_selfplay
        assert ref_mcts.__file__.startswith(REFERENCE) and ref_selfplay.__file__.startswith(REFERENCE)
        cfg = ref_mcts.Config()
        cfg.simulation_num = 32
        np.random.seed(2)
        rec = ref_selfplay.run(FakeNetwork(), ref_mcts.MCTS(cfg), max_moves=16)
        _check_record(rec)
        assert rec.ply > 4
    finally:
        sys.path[:] = saved_path
        for name in ("mcts", "selfplay", "network", "gamerecord", "minishogilib"):
            sys.modules.pop(name, None)
        sys.modules.update({k: v for k, v in saved_mods.items() if k in ("mcts", "selfplay", "network", "gamerecord", "minishogilib")})
