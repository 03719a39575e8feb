"""GPU-resident self-play: per-round kernel times (one cohort, serialised) and the two-cohort timeline."""
import os, sys
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network
games = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
plies = int(sys.argv[2]) if len(sys.argv) > 2 else 6
net = Network(precision="bf16", weights=W.init_weights(seed=55))
net.gpu_selfplay(games=games, total_moves=games, simulations=800, seed=77)
stats, _ = net.gpu_selfplay(games=games, total_moves=games * plies, simulations=800, mate_depth=7, seed=5)
print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in stats.items()})
