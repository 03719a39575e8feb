"""CPU oracle for the search tree the native engine implements (csrc/mcts.hpp).  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this file.  The reference drives ``minishogilib.MCTS`` (not vendored; ``mcts.py:27-116``) as
``set_root -> evaluate(root) -> [select_leaf -> evaluate -> backpropagate]*``; what that library computes inside is the
published AlphaZero search, restated here in the plainest form: PUCT selection ``Q + c_puct P sqrt(sum N) / (1 + N)``
with virtual losses counted as visits worth nothing, priors = softmax of the network's logits over the legal moves,
values stored for the side to move and flipped at every ply on the way up, terminal nodes from the game's rules in
selfplay.py's order (``selfplay.py:17-31``).  PARITY UNPINNED against minishogilib (its constants and tie-breaking are
not visible in the reference tree); the file pins the native tree to this restatement: ``tests/test_mcts_oracle.py``
runs both on the same positions, logits and values -- arithmetic in float32, operation for operation -- and requires
the same path for every selection and the same visit counts at the end.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


class Node:
    def __init__(self, parent=None, parent_edge=-1):
        self.parent, self.parent_edge = parent, parent_edge
        self.moves, self.prior, self.w, self.n_edge, self.vloss, self.child = [], [], [], [], [], []
        self.n = 0
        self.value = f32(0)
        self.expanded = self.terminal = False


class Tree:
    def __init__(self, c_puct=1.25):
        self.c_puct = f32(c_puct)
        self.root = Node()

    def select_leaf(self, node, pos):
        """PUCT descent from `node`; the moves are played on `pos` (an oracle Position); virtual loss on every edge taken."""
        while node.expanded and not node.terminal:
            total = sum(n + v for n, v in zip(node.n_edge, node.vloss))
            sq = np.sqrt(f32(total) + f32(1e-8))
            best, best_score = 0, f32(-1e30)
            for i in range(len(node.moves)):
                n = node.n_edge[i] + node.vloss[i]
                q = node.w[i] / f32(n) if n > 0 else f32(0)
                score = q + self.c_puct * node.prior[i] * sq / (f32(1) + f32(n))
                if score > best_score:
                    best, best_score = i, score
            node.vloss[best] += 1
            pos.do_move(node.moves[best])
            if node.child[best] is None:
                node.child[best] = Node(node, best)
            node = node.child[best]
        return node

    def evaluate(self, node, pos, moves, logits, value01):
        """Terminal test, else expansion: `moves` = the legal moves in the order the edges are to have, `logits` = the
        network's logit of each of them, value01 = the network's value for the side to move in [0, 1]."""
        if node.expanded or node.terminal:
            return
        rep, my_check, op_check = pos.is_repetition()
        terminal = None
        if my_check:
            terminal = 0.0                                    # the perpetual checker (to move) loses
        elif op_check:
            terminal = 1.0
        elif rep:
            terminal = 1.0 if pos.side == 1 else 0.0          # fourth occurrence: the second player wins
        elif not moves:
            terminal = 0.0                                    # no legal move: mated
        if terminal is not None:
            node.terminal, node.value = True, f32(terminal)
            return
        lg = np.asarray(logits, dtype=f32)
        e = np.exp(lg - lg.max()).astype(f32)
        total = f32(0)
        for x in e:
            total = f32(total + x)
        inv = f32(1) / total
        node.moves = list(moves)
        node.prior = [f32(x * inv) for x in e]
        node.w = [f32(0)] * len(moves)
        node.n_edge = [0] * len(moves)
        node.vloss = [0] * len(moves)
        node.child = [None] * len(moves)
        node.value = f32(value01)
        node.expanded = True

    def backpropagate(self, leaf):
        node, v = leaf, leaf.value
        node.n += 1
        while node.parent is not None:
            v = f32(1) - v
            p, i = node.parent, node.parent_edge
            if p.vloss[i] > 0:
                p.vloss[i] -= 1
            p.n_edge[i] += 1
            p.w[i] = f32(p.w[i] + v)
            node = p
            node.n += 1

    def visits(self, node=None):
        node = node or self.root
        return {m: n for m, n in zip(node.moves, node.n_edge) if n > 0}
