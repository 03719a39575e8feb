"""The native search tree (csrc/mcts.hpp behind erweitern_55_b200.minishogi.MCTS) against the plain restatement in
oracle/mcts_oracle.py: the same positions, logits and values go into both, every selection must walk the same path
and the visit counts must agree exactly at the end -- with leaf batches, so that virtual losses are in play, with
logits that rule moves out, and through terminal positions."""
import zlib

import numpy as np
import pytest

from erweitern_55_b200.minishogi import MCTS, Move, Position
from oracle import minishogi_oracle as O
from oracle.mcts_oracle import Tree

START = "rbsgk/4p/5/P4/KGSBR b - 1"


def _fake_network(key, moves):
    """Deterministic 'network' whose float32 results are exact: logits 0 or -1e30 (softmax = uniform over the moves
    left in), values k / 8."""
    h = zlib.crc32(key.encode())
    keep = [((h >> (i % 24)) ^ i) & 3 != 0 for i in range(len(moves))]
    if not any(keep):
        keep[h % len(moves)] = True
    return [0.0 if k else -1e30 for k in keep], (h % 9) / 8.0


@pytest.mark.parametrize("sfen,simulations,batch,ends", [
    (START, 640, 1, False),
    (START, 640, 8, False),                                          # leaf batches: virtual loss spreads the selections
    (START, 800, 16, None),                                          # as in self-play: forced playouts, pruned targets
    ("2k2/5/2B2/5/2K2 b GSBRgsr2p 1", 480, 8, True),                 # many drops, mates within reach
    ("2k2/5/5/5/2K2 b R 1 moves R*3c 3a2a 3c2c 2a3a 2c3c 3a2a 3c2c 2a3a 2c3c 3a2a 3c2c", 480, 4, True),   # perpetual check two plies below the root
])
def test_native_tree_walks_and_counts_like_the_oracle(sfen, simulations, batch, ends):
    native_root_pos, oracle_root_pos = Position(), O.Position(sfen)
    native_root_pos.set_sfen(sfen)
    native, oracle = MCTS(0.05), Tree()
    root = native.set_root(native_root_pos, False)

    def evaluate(node_n, pos_n, node_o, pos_o):
        moves = [m.sfen() for m in pos_n.generate_moves()]
        policy = np.zeros(1725, dtype=np.float32)
        logits, value = ([], 0.0)
        if moves:
            logits, value = _fake_network(pos_o.board_sfen(), moves)
            for m, lg in zip(moves, logits):
                policy[pos_n.policy_index(Move(m))] = lg
        native.evaluate(node_n, pos_n, policy, value)
        oracle.evaluate(node_o, pos_o, moves, logits, value)

    forced = ends is None                                            # mcts.py:97: forced_playouts during self-play
    played = []                                                      # moves played at the root since `sfen`

    def oracle_position():
        pos = O.Position(sfen)
        for m in played:
            pos.do_move(m)
        return pos

    def search(n_sims):
        if not native.expanded(root):                                # mcts.py:56-60
            assert not (oracle.root.expanded or oracle.root.terminal)
            evaluate(root, native_root_pos.copy(True), oracle.root, oracle_position())
            native.backpropagate(root)
            oracle.backpropagate(oracle.root)
        done = terminals = 0
        while done < n_sims:
            leaves = []
            for _ in range(batch):
                pos_n, pos_o = native_root_pos.copy(True), oracle_position()
                leaf_n = native.select_leaf(root, pos_n, forced)
                leaf_o = oracle.select_leaf(oracle.root, pos_o, forced)
                assert pos_n.sfen().rsplit(" ", 1)[0] == pos_o.board_sfen() and pos_n.get_ply() == pos_o.ply, (done, pos_n.sfen(True))
                leaves.append((leaf_n, pos_n, leaf_o, pos_o))
            for leaf_n, pos_n, leaf_o, pos_o in leaves:
                assert native.expanded(leaf_n) == (leaf_o.expanded or leaf_o.terminal)   # (mcts.py:99: nothing to evaluate)
                if not (leaf_o.expanded or leaf_o.terminal):
                    evaluate(leaf_n, pos_n, leaf_o, pos_o)
                terminals += leaf_o.terminal
                native.backpropagate(leaf_n)
                oracle.backpropagate(leaf_o)
                done += 1
        return terminals

    terminals = search(simulations)
    sum_n, q, visits = native.dump(root)
    assert dict(visits) == oracle.visits() and sum_n == simulations
    w = sum(float(x) for x in oracle.root.w)
    assert abs(q - w / simulations) < 1e-5
    if ends:
        assert terminals > 0                                         # the walk did run into ends of the game
    if forced:                                                       # the training targets (mcts.py:187, selfplay.py:80-92)
        pruned_sum, _, pruned = native.dump(root, True, True)
        assert dict(pruned) == oracle.visits(target_pruning=True) and pruned_sum == sum(dict(pruned).values()) <= simulations
        assert dict(pruned) != dict(visits)                          # (the pruning did take playouts back)

    # the game goes on: the most visited move is played and its subtree becomes the tree (set_root with reuse_tree,
    # mcts.py:50 / client.py:32) -- twice, the searches in between still in step
    for _ in range(2):
        if oracle.root.terminal or not oracle.root.moves:
            break
        best = max(range(len(oracle.root.moves)), key=lambda i: (oracle.root.n_edge[i], -i))
        move = oracle.root.moves[best]
        assert native.best_move(root).sfen() == move
        kept = oracle.root.child[best]
        played.append(move)
        native_root_pos.do_move(Move(move))
        assert native.set_root(native_root_pos, True) == root
        oracle.root = kept if kept is not None else Tree().root
        oracle.root.parent, oracle.root.parent_edge = None, -1
        if oracle.root.terminal:
            break
        kept_visits = sum(oracle.root.n_edge)
        assert native.get_playouts(root, True) == kept_visits
        search(160)
        sum_n, q, visits = native.dump(root)
        assert dict(visits) == oracle.visits() and sum_n == kept_visits + 160
