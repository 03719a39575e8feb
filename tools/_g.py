import sys, time
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network
net = Network(precision="bf16", weights=W.init_weights(seed=55))
for seed, moves in ((78, 16384), (6, 131072)):
    t0 = time.time()
    stats, _ = net.gpu_selfplay(games=16384, total_moves=moves, simulations=800, mate_depth=7, seed=seed)
    print(f"seed={seed}: {stats['moves_per_s']:.0f} moves/s, {stats['seconds']:.1f} s (wall {time.time()-t0:.1f}), {stats['games']} games", flush=True)
