"""Short ncu target for the GPU-resident self-play: a few rounds of 4096 games."""
import sys
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network
net = Network(precision="bf16", weights=W.init_weights(seed=55))
stats, _ = net.gpu_selfplay(games=4096, total_moves=4096, simulations=800, seed=5)
print("done", stats["moves"], stats["evals"])
