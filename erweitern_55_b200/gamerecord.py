"""Game record of one self-play game, field for field what the reference pickles to its trainer
(``erweitern_55/gamerecord.py:1-18``; sent by ``client.py:64-68``, consumed by ``train.py``)."""
from __future__ import annotations


class GameRecord:
    def __init__(self):
        self.ply = 0
        self.sfen_kif = []                 # moves as SFEN strings
        self.mcts_result = []              # per ply: (sum_N, Q, [(move, N), ...])
        self.learning_target_plys = []
        self.winner = 2                    # 0 first player, 1 second player, 2 undecided / draw
        self.timestamp = 0

    def to_dict(self):
        return {key: getattr(self, key)
                for key in ("ply", "sfen_kif", "mcts_result", "learning_target_plys", "winner", "timestamp")}

    @classmethod
    def from_native(cls, text: str, timestamp: int = 0):
        """Parse the text form ``e55_selfplay_run_ex`` hands to its sink (include/e55.h, ``e55_game_sink``)."""
        lines = text.strip().split("\n")
        head = lines[0].split()
        rec = cls()
        rec.winner = int(head[1])
        for line in lines[1:]:
            f = line.split()
            k = int(f[3])
            target = not f[0].startswith("~")       # fast searches under playout-cap oscillation (selfplay.py:90-91)
            rec.sfen_kif.append(f[0].lstrip("~"))
            rec.mcts_result.append((int(f[1]), float(f[2]), [(f[4 + 2 * i], int(f[5 + 2 * i])) for i in range(k)]))
            if target:
                rec.learning_target_plys.append(rec.ply)
            rec.ply += 1
        rec.timestamp = timestamp
        return rec
