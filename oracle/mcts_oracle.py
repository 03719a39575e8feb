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
        self.forced_k = f32(2)
        self.root = Node()

    def select_leaf(self, node, pos, forced_playouts=False):
        """PUCT descent from `node`; the moves are played on `pos` (an oracle Position); virtual loss on every edge taken.
        forced_playouts (KataGo, used by mcts.py:97 during self-play): a child of the starting node that has been
        visited is taken again while it has fewer than sqrt(k P N) playouts, k = 2."""
        start = node
        while node.expanded and not node.terminal:
            total = sum(n + v for n, v in zip(node.n_edge, node.vloss))
            sq = np.sqrt(f32(total) + f32(1e-8))
            best, best_score = None, f32(-1e30)
            if forced_playouts and node is start:
                for i in range(len(node.moves)):
                    if node.n_edge[i] > 0 and f32(node.n_edge[i] + node.vloss[i]) < np.sqrt(self.forced_k * node.prior[i] * f32(total)):
                        best = i
                        break
            for i in range(len(node.moves) if best is None else 0):
                n = node.n_edge[i] + node.vloss[i]
                q = node.w[i] / f32(n) if n > 0 else f32(0)
                score = q + self.c_puct * node.prior[i] * sq / (f32(1) + f32(n))
                if score > best_score:
                    best, best_score = i, score
            best = best or 0
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

    def visits(self, node=None, target_pruning=False):
        """Visit counts of the children.  target_pruning (KataGo; mcts.py:187 for the training targets): forced playouts
        are taken back from every child but the most visited one while its PUCT score, with its utility held, stays
        below the best child's; children left with one playout are dropped."""
        node = node or self.root
        cnt = list(node.n_edge)
        total = sum(cnt)
        if target_pruning and total > 0:
            best = max(range(len(cnt)), key=lambda i: (cnt[i], -i))
            sq = np.sqrt(f32(total))

            def puct(i, n):
                q = node.w[i] / f32(node.n_edge[i]) if node.n_edge[i] > 0 else f32(0)
                return q + self.c_puct * node.prior[i] * sq / (f32(1) + f32(n))
            best_score = puct(best, cnt[best])
            for i in range(len(cnt)):
                if i == best or cnt[i] == 0:
                    continue
                forced = int(np.sqrt(self.forced_k * node.prior[i] * f32(total)))
                while forced > 0 and cnt[i] > 0 and puct(i, cnt[i] - 1) < best_score:
                    cnt[i] -= 1
                    forced -= 1
                if cnt[i] <= 1:
                    cnt[i] = 0
        return {m: n for m, n in zip(node.moves, cnt) if n > 0}
