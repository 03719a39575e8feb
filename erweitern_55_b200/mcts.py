"""Search driver with the reference's surface: ``Config`` and ``MCTS`` of ``erweitern_55/mcts.py:8-205``.

``MCTS.run(position, nn, timelimit, verbose, set_root_only)`` keeps the reference's contract (root set with optional
tree reuse, root evaluation, optional Dirichlet noise, leaf batches of ``config.batch_size`` until
``config.simulation_num`` simulations, the time limit, a memory usage above 90 % or ``stop()`` end the search; USI
``info`` lines every 50 batches when verbose) and returns the root node.  When ``nn`` is this package's ``Network`` a
batch is one native call (``e55_mcts_search_batches``: select, bit-packed encode, ONE forward, expand, back up);
any other object with the reference's ``get_input / get_inputs / predict`` methods is driven leaf by leaf from
Python exactly as the reference does (mcts.py:91-110).
"""
from __future__ import annotations

import threading
import time

from . import minishogi


class Config:
    def __init__(self):                    # defaults of mcts.py:9-21
        self.memory_size = 0.2             # GB
        self.batch_size = 16
        self.simulation_num = 800
        self.use_dirichlet = False
        self.forced_playouts = False
        self.reuse_tree = True
        self.target_pruning = False
        self.immediate = False


class MCTS:
    def __init__(self, config):
        self.config = config
        self.mcts = minishogi.MCTS(config.memory_size)
        self.searching = False
        self.lock = threading.Lock()

    def clear(self):
        self.mcts.clear()

    # -- helpers ---------------------------------------------------------------------------------------------
    @staticmethod
    def _is_native(nn):
        return hasattr(nn, "_ctx") and hasattr(nn, "predict_packed")

    def _log(self, root):
        pv_moves, q = self.mcts.info(root)
        print("info depth {} nodes {} hashfull {} score winrate {:.3f} pv {}".format(
            len(pv_moves), self.mcts.get_nodes(), int(self.mcts.get_usage() * 1000), q,
            " ".join(m.sfen() for m in pv_moves)), flush=True)

    def _python_batch(self, root, position, nn):
        cfg = self.config
        positions = [position.copy(True) for _ in range(cfg.batch_size)]
        leaves = [self.mcts.select_leaf(root, p, cfg.forced_playouts) for p in positions]
        policy, value = nn.predict(nn.get_inputs(positions))
        value = (value + 1) / 2
        for b, leaf in enumerate(leaves):
            self.mcts.evaluate(leaf, positions[b], policy[b], value[b][0])
        for leaf in leaves:
            self.mcts.backpropagate(leaf)

    # -- the search ------------------------------------------------------------------------------------------
    def run(self, position, nn, timelimit=0, verbose=False, set_root_only=False):
        """Search ``position``; ``timelimit`` in milliseconds (0 = none).  Returns the root node."""
        cfg = self.config
        self.searching = True
        started = time.time()
        native = self._is_native(nn)
        use_service = bool(getattr(nn, "_service_cap", 0))

        if set_root_only:
            self.mcts.set_root(position, cfg.reuse_tree)
            return None
        if native:
            terminal = self.mcts.search_root(position, nn, cfg.reuse_tree, cfg.use_dirichlet, use_service)
            root = 0
        else:
            root = self.mcts.set_root(position, cfg.reuse_tree)
            if not self.mcts.expanded(root):
                policy, value = nn.predict(nn.get_input(position, True))
                value = (value + 1) / 2
                self.mcts.evaluate(root, position, policy[0], value[0][0])
            if cfg.use_dirichlet:
                self.mcts.add_noise(root)
            terminal = self.mcts.best_move(root) is None
        if terminal:
            return root

        batches = 0
        simulations = 0
        while simulations < cfg.simulation_num:
            if self.mcts.get_usage() > 0.9:
                break
            with self.lock:
                if not self.searching:
                    break
            if timelimit > 0 and (time.time() - started) * 1000 >= timelimit:
                break
            if cfg.immediate and self.mcts.get_playouts(root, True) >= cfg.simulation_num:
                return root
            if native:
                self.mcts.search_batches(position, nn, 1, cfg.batch_size, cfg.forced_playouts, use_service)
            else:
                self._python_batch(root, position, nn)
            if verbose and batches % 50 == 0:
                self._log(root)
            batches += 1
            simulations += cfg.batch_size
        if verbose:
            self._log(root)
        return root

    def stop(self):
        with self.lock:
            self.searching = False

    # -- results (mcts.py:140-205) ------------------------------------------------------------------------------
    def best_move(self, node):
        return self.mcts.best_move(node)

    def softmax_sample(self, node, temperature):
        return self.mcts.softmax_sample(node, temperature)

    def softmax_sample_among_top_moves(self, node, away=0.01, temperature=10.0):
        return self.mcts.softmax_sample_among_top_moves(node, away, temperature)

    def dump(self, node, remove_zeros=True):
        return self.mcts.dump(node, self.config.target_pruning, remove_zeros)

    def print(self, node):
        self.mcts.print(node)

    def debug(self, node):
        self.mcts.debug(node)

    def visualize(self, node, node_num=20):
        return self.mcts.visualize(node, node_num)
