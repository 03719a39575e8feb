"""CPU oracle for the minishogi rules the native engine implements.  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this file; the product (``erweitern_55_b200``: ``csrc/minishogi.hpp``, the bitboard engine
behind ``Position``) never does.

What it restates.  The reference takes its rules from ``minishogilib`` (``requirements.txt:16``, a git dependency that
is not vendored in ``/root/reference``), through ``Position.set_sfen / do_move / generate_moves / is_repetition /
solve_checkmate_dfs`` (``selfplay.py:11-36,78``; ``usi.py:44-140``).  The algorithm restated is therefore the published
game: 5x5 board, K G S B R P per side, promotion on the far rank only (a pawn must), drops on empty squares with the
two-pawn, last-rank and pawn-drop-mate restrictions, fourth occurrence of a position = repetition, perpetual check
loses.  PINNED on the reference's own fixtures -- the 9 known answers of ``tests/test_checkmate.py`` and the 8 of
``tests/test_repetition.py`` -- and on the published perft counts of the start position (14, 181, 2 512, 35 401);
``tests/test_minishogi_oracle.py`` checks those first and then holds the native engine against this file on random
positions.  ``to_alphazero_input`` / ``policy_index`` restate this repo's own plane and move layout (minishogilib's is
not visible in the reference tree: "parity unpinned" for those two); they keep the native encoders honest to their
documentation.

Deliberately naive and independent of the engine's design: a mailbox board, pseudo-legal moves from step tables, legality
by playing the move and looking for an attacker of the king.  Slow (a few thousand positions per second) and small.
"""
from __future__ import annotations

BLACK, WHITE = 0, 1          # BLACK = first player, upper-case letters, moves towards rank 'a' (row 0)
N = 5
PROMOTABLE = {"S": "+S", "B": "+B", "R": "+R", "P": "+P"}
UNPROMOTE = {v: k for k, v in PROMOTABLE.items()}
HAND_ORDER = "GSBRP"

# one-step directions as (d_row, d_col) for BLACK; WHITE mirrors the row sign
_GOLD = [(-1, 0), (-1, 1), (-1, -1), (0, 1), (0, -1), (1, 0)]
STEPS = {
    "K": [(-1, 0), (-1, 1), (-1, -1), (0, 1), (0, -1), (1, 0), (1, 1), (1, -1)],
    "G": _GOLD, "+S": _GOLD, "+P": _GOLD,
    "S": [(-1, 0), (-1, 1), (-1, -1), (1, 1), (1, -1)],
    "P": [(-1, 0)],
    "+B": [(-1, 0), (1, 0), (0, 1), (0, -1)],
    "+R": [(-1, 1), (-1, -1), (1, 1), (1, -1)],
}
SLIDES = {
    "B": [(-1, 1), (-1, -1), (1, 1), (1, -1)], "+B": [(-1, 1), (-1, -1), (1, 1), (1, -1)],
    "R": [(-1, 0), (1, 0), (0, 1), (0, -1)], "+R": [(-1, 0), (1, 0), (0, 1), (0, -1)],
}


def _square_name(r, c):
    return "54321"[c] + "abcde"[r]


def _parse_square(s):
    return "abcde".index(s[1]), "54321".index(s[0])


class Position:
    """board[r][c] = None or (colour, kind) with kind in K G S B R P +S +B +R +P; hands[colour][kind] = count."""

    def __init__(self, sfen="rbsgk/4p/5/P4/KGSBR b - 1"):
        self.set_sfen(sfen)

    # ---- SFEN ------------------------------------------------------------------------------------------------------
    def set_sfen(self, sfen):
        parts = sfen.split()
        self.board = [[None] * N for _ in range(N)]
        for r, row in enumerate(parts[0].split("/")):
            c, promoted = 0, False
            for ch in row:
                if ch == "+":
                    promoted = True
                elif ch.isdigit():
                    c += int(ch)
                else:
                    kind = ch.upper()
                    self.board[r][c] = (BLACK if ch.isupper() else WHITE, "+" + kind if promoted else kind)
                    promoted = False
                    c += 1
        self.side = WHITE if parts[1] == "w" else BLACK
        self.hands = [dict.fromkeys(HAND_ORDER, 0), dict.fromkeys(HAND_ORDER, 0)]
        count = ""
        for ch in parts[2]:
            if ch == "-":
                continue
            if ch.isdigit():
                count += ch
                continue
            self.hands[BLACK if ch.isupper() else WHITE][ch.upper()] += int(count) if count else 1
            count = ""
        self.ply = 0
        # history of (position key, the move into it gave check); entry 0 = the position set here
        self.history = [(self.key(), False)]
        self.snapshots = [self._snapshot()]
        if "moves" in parts:
            for m in parts[parts.index("moves") + 1:]:
                self.do_move(m)

    def board_sfen(self):
        rows = []
        for r in range(N):
            s, empty = "", 0
            for c in range(N):
                p = self.board[r][c]
                if p is None:
                    empty += 1
                    continue
                if empty:
                    s += str(empty)
                    empty = 0
                letter = p[1][-1] if p[0] == BLACK else p[1][-1].lower()
                s += ("+" if p[1][0] == "+" else "") + letter
            rows.append(s + (str(empty) if empty else ""))
        hand = ""
        for colour in (BLACK, WHITE):
            for k in HAND_ORDER:
                n = self.hands[colour][k]
                if n:
                    hand += (str(n) if n > 1 else "") + (k if colour == BLACK else k.lower())
        return "/".join(rows) + (" b " if self.side == BLACK else " w ") + (hand or "-")

    def key(self):
        return self.board_sfen()

    # ---- attacks ---------------------------------------------------------------------------------------------------
    def _reach(self, r, c, colour, kind):
        """Squares the piece on (r, c) attacks (own pieces included as targets are filtered by the caller)."""
        sign = 1 if colour == BLACK else -1
        for dr, dc in STEPS.get(kind, ()):
            rr, cc = r + sign * dr, c + sign * dc
            if 0 <= rr < N and 0 <= cc < N:
                yield rr, cc
        for dr, dc in SLIDES.get(kind, ()):
            rr, cc = r + dr, c + dc
            while 0 <= rr < N and 0 <= cc < N:
                yield rr, cc
                if self.board[rr][cc] is not None:
                    break
                rr, cc = rr + dr, cc + dc

    def attacked(self, r, c, by):
        for rr in range(N):
            for cc in range(N):
                p = self.board[rr][cc]
                if p is not None and p[0] == by and (r, c) in self._reach(rr, cc, p[0], p[1]):
                    return True
        return False

    def king_square(self, colour):
        for r in range(N):
            for c in range(N):
                if self.board[r][c] == (colour, "K"):
                    return r, c
        return None

    def in_check(self, colour):
        k = self.king_square(colour)
        return k is not None and self.attacked(k[0], k[1], 1 - colour)

    # ---- moves -----------------------------------------------------------------------------------------------------
    def _pseudo_moves(self):
        me = self.side
        far = 0 if me == BLACK else N - 1
        for r in range(N):
            for c in range(N):
                p = self.board[r][c]
                if p is None or p[0] != me:
                    continue
                for rr, cc in self._reach(r, c, me, p[1]):
                    t = self.board[rr][cc]
                    if t is not None and t[0] == me:
                        continue
                    name = _square_name(r, c) + _square_name(rr, cc)
                    can_promote = p[1] in PROMOTABLE and (r == far or rr == far)
                    if can_promote:
                        yield name + "+"
                    if not (p[1] == "P" and rr == far):             # a pawn on the far rank could never move again
                        yield name
        for k in HAND_ORDER:
            if not self.hands[me][k]:
                continue
            for r in range(N):
                for c in range(N):
                    if self.board[r][c] is not None:
                        continue
                    if k == "P":
                        if r == far:
                            continue
                        if any(self.board[x][c] == (me, "P") for x in range(N)):   # two pawns on a file
                            continue
                    yield k + "*" + _square_name(r, c)

    def _apply(self, move):
        """Plays a pseudo-legal move; returns what is needed to take it back."""
        me = self.side
        if move[1] == "*":
            r, c = _parse_square(move[2:4])
            self.board[r][c] = (me, move[0])
            self.hands[me][move[0]] -= 1
            undo = ("drop", r, c, move[0])
        else:
            r, c = _parse_square(move[0:2])
            rr, cc = _parse_square(move[2:4])
            piece, target = self.board[r][c], self.board[rr][cc]
            if target is not None:
                self.hands[me][UNPROMOTE.get(target[1], target[1])] += 1
            self.board[rr][cc] = (me, PROMOTABLE[piece[1]]) if move.endswith("+") else piece
            self.board[r][c] = None
            undo = ("move", r, c, rr, cc, piece, target)
        self.side = 1 - me
        return undo

    def _revert(self, undo):
        self.side = 1 - self.side
        me = self.side
        if undo[0] == "drop":
            _, r, c, k = undo
            self.board[r][c] = None
            self.hands[me][k] += 1
        else:
            _, r, c, rr, cc, piece, target = undo
            self.board[r][c] = piece
            self.board[rr][cc] = target
            if target is not None:
                self.hands[me][UNPROMOTE.get(target[1], target[1])] -= 1

    def generate_moves(self):
        """Fully legal moves of the side to move, as USI strings."""
        me = self.side
        legal = []
        for move in list(self._pseudo_moves()):
            undo = self._apply(move)
            ok = not self.in_check(me)
            if ok and move[0] == "P" and move[1] == "*" and self.in_check(1 - me):
                ok = bool(self._has_legal_move())                   # a pawn drop may check, but not mate
            self._revert(undo)
            if ok:
                legal.append(move)
        return legal

    def _has_legal_move(self):
        me = self.side
        for move in list(self._pseudo_moves()):
            undo = self._apply(move)
            ok = not self.in_check(me)
            if ok and move[0] == "P" and move[1] == "*" and self.in_check(1 - me):
                ok = bool(self._has_legal_move())
            self._revert(undo)
            if ok:
                return True
        return False

    def do_move(self, move):
        self._apply(move)
        self.ply += 1
        self.history.append((self.key(), self.in_check(self.side)))
        self.snapshots.append(self._snapshot())

    def undo_move(self, undo):
        self.history.pop()
        self.snapshots.pop()
        self.ply -= 1
        self._revert(undo)

    # ---- network input and policy index (this repo's own layout, documented in csrc/minishogi.hpp) ---------------------
    def _snapshot(self):
        key = self.history[-1][0]
        return ([row[:] for row in self.board], [dict(h) for h in self.hands], sum(1 for k, _ in self.history if k == key))

    def to_alphazero_input(self):
        """[266][5][5] floats: 8 history steps (newest first) x 33 planes -- 10 piece kinds of the side to move, 10 of the
        other side, 3 one-hot planes for the 1st / 2nd / 3rd+ occurrence, 5 + 5 hand counts / 2 -- then a plane of ones
        if the second player is to move and a plane of ply / 512.  Everything is seen from the side to move: for the
        second player the board is turned by 180 degrees."""
        kinds = ["K", "G", "S", "B", "R", "P", "+S", "+B", "+R", "+P"]
        me = self.side
        x = [[[0.0] * N for _ in range(N)] for _ in range(266)]

        def fill(plane, v):
            for r in range(N):
                for c in range(N):
                    x[plane][r][c] = v
        for t in range(8):
            if t >= len(self.snapshots):
                break
            board, hands, occurrences = self.snapshots[-1 - t]
            base = 33 * t
            for r in range(N):
                for c in range(N):
                    piece = board[r][c]
                    if piece is None:
                        continue
                    rr, cc = (r, c) if me == BLACK else (N - 1 - r, N - 1 - c)
                    x[base + (0 if piece[0] == me else 10) + kinds.index(piece[1])][rr][cc] = 1.0
            fill(base + 20 + min(occurrences, 3) - 1, 1.0)
            for k, kind in enumerate(HAND_ORDER):
                if hands[me][kind]:
                    fill(base + 23 + k, hands[me][kind] / 2.0)
                if hands[1 - me][kind]:
                    fill(base + 28 + k, hands[1 - me][kind] / 2.0)
        if me == WHITE:
            fill(264, 1.0)
        if self.ply > 0:
            fill(265, min(self.ply, 512) / 512.0)
        return x

    def policy_index(self, move):
        """Index of a move in the 69 x 25 policy: planes 0-31 = direction (N NE E SE S SW W NW, seen from the side to
        move) x distance 1-4 at the origin square, 32-63 the same with promotion, 64-68 drops of G S B R P at the
        destination square."""
        def persp(r, c):
            return (r, c) if self.side == BLACK else (N - 1 - r, N - 1 - c)
        if move[1] == "*":
            r, c = persp(*_parse_square(move[2:4]))
            return (64 + HAND_ORDER.index(move[0])) * 25 + r * N + c
        r, c = persp(*_parse_square(move[0:2]))
        rr, cc = persp(*_parse_square(move[2:4]))
        dr, dc = rr - r, cc - c
        step = ((dr > 0) - (dr < 0), (dc > 0) - (dc < 0))
        d = [(-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1)].index(step)
        dist = max(abs(dr), abs(dc))
        return ((32 if move.endswith("+") else 0) + d * 4 + dist - 1) * 25 + r * N + c

    # ---- repetition (tests/test_repetition.py) ---------------------------------------------------------------------
    def is_repetition(self):
        """(repetition, the side to move did the perpetual check, the other side did)."""
        key = self.history[-1][0]
        same = [i for i, (k, _) in enumerate(self.history) if k == key]
        if len(same) < 4:
            return (False, False, False)
        last = len(self.history) - 1
        mine = [self.history[i][1] for i in range(same[0] + 1, last + 1) if (last - i) % 2 == 1]
        theirs = [self.history[i][1] for i in range(same[0] + 1, last + 1) if (last - i) % 2 == 0]
        return (True, bool(mine) and all(mine), bool(theirs) and all(theirs))

    # ---- mate search (tests/test_checkmate.py) ---------------------------------------------------------------------
    def solve_checkmate_dfs(self, depth):
        """(found, move): can the side to move force mate within `depth` plies giving check with every move?"""
        return self._attack(depth)

    def _attack(self, depth):
        if depth <= 0:
            return (False, None)
        me = self.side
        for move in self.generate_moves():
            undo = self._apply(move)
            mate = self.in_check(1 - me) and self._defend(depth - 1)
            self._revert(undo)
            if mate:
                return (True, move)
        return (False, None)

    def _defend(self, depth):
        """True if every reply of the side to move (which is in check) still loses within `depth` plies."""
        replies = self.generate_moves()
        if not replies:
            return True
        if depth <= 1:
            return False
        for move in replies:
            undo = self._apply(move)
            found, _ = self._attack(depth - 1)
            self._revert(undo)
            if not found:
                return False
        return True


def perft(pos, depth):
    moves = pos.generate_moves()
    if depth <= 1:
        return len(moves)
    total = 0
    for m in moves:
        undo = pos._apply(m)
        total += perft(pos, depth - 1)
        pos._revert(undo)
    return total
