"""Short ncu target for the self-play path: a few hundred plies of real self-play on the evaluation service
(one tower_kernel per coalesced batch: the legal-move softmax is its policy tail)."""
import sys
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network

net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=1024)
net.start_service(max_batch=1024, max_wait_us=200)
st = net.selfplay(threads=8, total_moves=int(sys.argv[1]) if len(sys.argv) > 1 else 300, simulations=800, games_per_thread=16)
net.stop_service()
print("done", st["moves"], st["evals"], st["mean_gpu_batch"])
