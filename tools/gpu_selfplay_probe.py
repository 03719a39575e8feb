"""GPU-resident self-play: device perft, a validated sample of games, throughput vs number of games."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
from erweitern_55_b200 import _capi, weights as W
from erweitern_55_b200.minishogi import Move, Position
from erweitern_55_b200.network import Network
lib = _capi.load()
for d in range(1, 6):
    n = C.c_int64()
    rc = lib.e55_selftest_engine(0, d, C.byref(n)); print("device perft", d, n.value, "rc", rc, flush=True)
net = Network(precision="bf16", weights=W.init_weights(seed=55))
stats, games = net.gpu_selfplay(games=256, total_moves=3000, simulations=64, records=64, seed=3)
print(stats, len(games), flush=True)
bad = 0
for rec in games:
    p = Position()
    for m in rec.sfen_kif:
        if m not in {x.sfen() for x in p.generate_moves()}:
            bad += 1; print("ILLEGAL", m, "in", p.sfen()); break
        p.do_move(Move(m))
    else:
        if not ((not p.generate_moves()) or p.is_repetition()[0] or rec.ply >= 512):
            bad += 1; print("game not over?", p.sfen(), rec.winner, rec.ply)
print("validated", len(games), "games, problems:", bad, flush=True)
for g in (int(a) for a in sys.argv[1:]):
    stats, _ = net.gpu_selfplay(games=g, total_moves=max(2000, g * 4), simulations=800, seed=5)
    print(f"games={g}: {stats['moves_per_s']:.0f} moves/s, {stats['evals_per_s']/1e6:.2f} M useful evals/s, {stats['seconds']:.1f} s, "
          f"{stats['games']} games finished ({stats['mean_game_plies']:.1f} plies)", flush=True)
