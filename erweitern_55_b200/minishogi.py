"""Python face of the native minishogi engine (``csrc/minishogi.hpp``), shaped like the parts of
``minishogilib`` the reference calls: ``Position`` (``selfplay.py:11-36,78``, ``usi.py``, the reference's
``tests/test_checkmate.py`` / ``tests/test_repetition.py``) and ``Move``.  minishogilib itself (Rust, v0.6.12)
is not vendored by the reference and cannot be built here; rules, SFEN handling, repetition and the mate
search are pinned by the reference's 17 known-answer tests, the input encoding and the policy indexing are
this repo's own restatement (see the header comments of ``minishogi.hpp``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .packing import unpack_planes

IN_PLANES = 266


__version__ = "0.6.12+e55b200"   # the minishogilib release the reference pins (.github/workflows/test.yml:11-12)


class Move:
    def __init__(self, sfen: str):
        self._sfen = sfen

    def __str__(self):                     # usi.py:132 prints 'bestmove {}'.format(move)
        return self._sfen

    def sfen(self) -> str:
        return self._sfen

    def __eq__(self, other):
        return isinstance(other, Move) and other._sfen == self._sfen

    def __hash__(self):
        return hash(self._sfen)

    def __repr__(self):
        return f"Move({self._sfen})"


class Position:
    def __init__(self, _handle=None):
        self._lib = _capi.load()
        self._h = _handle if _handle is not None else self._lib.e55_pos_create()
        if not self._h:
            raise MemoryError("e55_pos_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.e55_pos_destroy(h)

    def set_start_position(self):
        self.set_sfen("rbsgk/4p/5/P4/KGSBR b - 1")

    def set_sfen(self, sfen: str):
        if self._lib.e55_pos_set_sfen(self._h, sfen.encode()) != 0:
            raise ValueError(f"bad SFEN: {sfen!r}")

    def sfen(self) -> str:
        buf = C.create_string_buffer(256)
        self._lib.e55_pos_sfen(self._h, buf, 256)
        return buf.value.decode()

    def copy(self, _with_history=True):
        return Position(self._lib.e55_pos_copy(self._h))

    def get_side_to_move(self) -> int:
        return self._lib.e55_pos_side_to_move(self._h)

    def get_ply(self) -> int:
        return self._lib.e55_pos_ply(self._h)

    def generate_moves(self):
        buf = C.create_string_buffer(4096)
        n = self._lib.e55_pos_generate_moves(self._h, buf, 4096)
        if n < 0:
            raise RuntimeError("generate_moves failed")
        return [Move(s) for s in buf.value.decode().split()]

    def do_move(self, move):
        s = move.sfen() if isinstance(move, Move) else str(move)
        if self._lib.e55_pos_do_move(self._h, s.encode()) != 0:
            raise ValueError(f"illegal move {s}")

    def is_repetition(self):
        out = (C.c_int * 3)()
        self._lib.e55_pos_is_repetition(self._h, out)
        return bool(out[0]), bool(out[1]), bool(out[2])

    def solve_checkmate_dfs(self, depth: int):
        buf = C.create_string_buffer(16)
        r = self._lib.e55_pos_solve_checkmate_dfs(self._h, depth, buf, 16)
        if r < 0:
            raise RuntimeError("solve_checkmate_dfs failed")
        return (True, Move(buf.value.decode())) if r else (False, None)

    def pretty(self) -> str:
        buf = C.create_string_buffer(512)
        self._lib.e55_pos_print(self._h, buf, 512)
        return buf.value.decode()

    def print(self):
        print(self.pretty(), end="", flush=True)

    def __str__(self):
        return self.sfen()

    def to_packed_input(self):
        bits = np.zeros(IN_PLANES, dtype=np.uint32)
        scale = np.zeros(IN_PLANES, dtype=np.float32)
        self._lib.e55_pos_encode_packed(self._h, bits.ctypes.data, scale.ctypes.data)
        return bits, scale

    def to_alphazero_input(self):
        """Flat 266*25 float32 sequence, as ``Network.get_input`` reshapes it (``network.py:255-258``)."""
        bits, scale = self.to_packed_input()
        return unpack_planes(bits[None], scale[None]).reshape(-1)

    def policy_index(self, move) -> int:
        s = move.sfen() if isinstance(move, Move) else str(move)
        r = self._lib.e55_pos_policy_index(self._h, s.encode())
        if r < 0:
            raise ValueError(f"illegal move {s}")
        return r


def encode_packed_batch(positions):
    """Packed network input of many positions in one native call: ``bits`` uint32 ``[N,266]``, ``scale`` float32."""
    lib = _capi.load()
    n = len(positions)
    bits = np.empty((n, IN_PLANES), dtype=np.uint32)
    scale = np.empty((n, IN_PLANES), dtype=np.float32)
    handles = (C.c_void_p * n)(*[p._h for p in positions])
    if lib.e55_pos_encode_packed_batch(handles, n, bits.ctypes.data, scale.ctypes.data) != 0:
        raise RuntimeError("e55_pos_encode_packed_batch failed")
    return bits, scale


class MCTS:
    """``minishogilib.MCTS`` as ``mcts.py`` drives it (mcts.py:27-205), on the native tree of ``csrc/mcts.hpp``.
    Nodes are plain integers, valid until the next ``set_root`` / ``clear``."""

    def __init__(self, memory_size=0.2):
        self._lib = _capi.load()
        self._h = self._lib.e55_mcts_create(float(memory_size))
        if not self._h:
            raise MemoryError("e55_mcts_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.e55_mcts_destroy(h)

    @staticmethod
    def _ok(rc, what):
        if rc < 0:
            raise RuntimeError(f"{what} failed ({rc})")
        return rc

    def configure(self, c_puct=1.25, dirichlet_alpha=0.15, dirichlet_eps=0.25, seed=0x55):
        self._ok(self._lib.e55_mcts_configure(self._h, c_puct, dirichlet_alpha, dirichlet_eps, seed), "configure")

    def clear(self):
        self._lib.e55_mcts_clear(self._h)

    def set_root(self, position, reuse_tree=True):
        return self._ok(self._lib.e55_mcts_set_root(self._h, position._h, 1 if reuse_tree else 0), "set_root")

    def expanded(self, node):
        return bool(self._ok(self._lib.e55_mcts_expanded(self._h, node), "expanded"))

    def evaluate(self, node, position, policy, value):
        pol = np.ascontiguousarray(policy, dtype=np.float32)
        if pol.shape != (1725,):
            raise ValueError("policy must hold 1725 logits")
        self._ok(self._lib.e55_mcts_evaluate(self._h, node, position._h, pol.ctypes.data, float(value)), "evaluate")

    def add_noise(self, node):
        self._ok(self._lib.e55_mcts_add_noise(self._h, node), "add_noise")

    def select_leaf(self, node, position, forced_playouts=False):
        return self._ok(self._lib.e55_mcts_select_leaf(self._h, node, position._h, 1 if forced_playouts else 0), "select_leaf")

    def backpropagate(self, leaf):
        self._ok(self._lib.e55_mcts_backpropagate(self._h, leaf), "backpropagate")

    def get_usage(self):
        return float(self._lib.e55_mcts_get_usage(self._h))

    def get_nodes(self):
        return self._lib.e55_mcts_get_nodes(self._h)

    def get_playouts(self, node, child_sum=False):
        return self._ok(self._lib.e55_mcts_get_playouts(self._h, node, 1 if child_sum else 0), "get_playouts")

    def _move(self, rc, buf, what):
        self._ok(rc, what)
        return Move(buf.value.decode()) if rc else None

    def best_move(self, node):
        buf = C.create_string_buffer(16)
        return self._move(self._lib.e55_mcts_best_move(self._h, node, buf, 16), buf, "best_move")

    def softmax_sample(self, node, temperature=1.0):
        buf = C.create_string_buffer(16)
        return self._move(self._lib.e55_mcts_softmax_sample(self._h, node, temperature, buf, 16), buf, "softmax_sample")

    def softmax_sample_among_top_moves(self, node, away=0.01, temperature=10.0):
        buf = C.create_string_buffer(16)
        return self._move(self._lib.e55_mcts_softmax_sample_among_top_moves(self._h, node, away, temperature, buf, 16),
                          buf, "softmax_sample_among_top_moves")

    def info(self, node):
        buf = C.create_string_buffer(4096)
        q = C.c_float()
        self._ok(self._lib.e55_mcts_info(self._h, node, buf, 4096, C.byref(q)), "info")
        return [Move(s) for s in buf.value.decode().split()], float(q.value)

    def dump(self, node, target_pruning=False, remove_zeros=True):
        """``(sum_N, Q, [(move_sfen, N), ...])`` -- the ``mcts_result`` entry of a game record (gamerecord.py:5)."""
        buf = C.create_string_buffer(4096)
        counts = (C.c_int * 256)()
        sum_n, q = C.c_int(), C.c_float()
        k = self._ok(self._lib.e55_mcts_dump(self._h, node, 1 if target_pruning else 0, 1 if remove_zeros else 0,
                                             C.byref(sum_n), C.byref(q), buf, 4096, counts, 256), "dump")
        moves = buf.value.decode().split()
        return int(sum_n.value), float(q.value), [(moves[i], int(counts[i])) for i in range(k)]

    def describe(self, node):
        buf = C.create_string_buffer(1 << 15)
        self._ok(self._lib.e55_mcts_describe(self._h, node, buf, 1 << 15), "describe")
        return buf.value.decode()

    def print(self, node):
        print(self.describe(node), end="", flush=True)

    debug = print

    def visualize(self, node, node_num=20):
        buf = C.create_string_buffer(1 << 16)
        self._ok(self._lib.e55_mcts_visualize(self._h, node, node_num, buf, 1 << 16), "visualize")
        return buf.value.decode()

    # ---- native halves of mcts.MCTS.run on a Network (no Python per leaf) ---------------------------------------
    def search_root(self, position, nn, reuse_tree=True, use_dirichlet=False, use_service=False):
        """Set the root, evaluate it if needed, add noise (mcts.py:49-64).  True when the root is terminal."""
        rc = self._lib.e55_mcts_search_root(self._h, nn._ctx, position._h, int(reuse_tree), int(use_dirichlet), int(use_service))
        if rc < 0:
            _capi.check(nn._ctx, rc)
        return rc == 1

    def search_batches(self, position, nn, n_batches, leaf_batch, forced_playouts=False, use_service=False):
        """``n_batches`` iterations of the select / evaluate / back-up loop (mcts.py:91-110); simulations done."""
        rc = self._lib.e55_mcts_search_batches(self._h, nn._ctx, position._h, n_batches, leaf_batch, int(forced_playouts),
                                               int(use_service))
        if rc < 0:
            _capi.check(nn._ctx, rc)
        return rc
