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


class Move:
    def __init__(self, sfen: str):
        self._sfen = sfen

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

    def to_packed_input(self):
        bits = np.zeros(IN_PLANES, dtype=np.uint32)
        scale = np.zeros(IN_PLANES, dtype=np.float32)
        self._lib.e55_pos_encode_packed(self._h, bits.ctypes.data, scale.ctypes.data)
        return bits, scale

    def to_alphazero_input(self):
        """Flat 266*25 float list, as ``Network.get_input`` reshapes it (``network.py:255-258``)."""
        bits, scale = self.to_packed_input()
        return unpack_planes(bits[None], scale[None]).reshape(-1).tolist()

    def policy_index(self, move) -> int:
        s = move.sfen() if isinstance(move, Move) else str(move)
        r = self._lib.e55_pos_policy_index(self._h, s.encode())
        if r < 0:
            raise ValueError(f"illegal move {s}")
        return r
