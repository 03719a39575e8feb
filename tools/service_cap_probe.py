"""Host self-play driver on the evaluation service: moves/s against the service's batch cap (1024 vs 1184 = 148 SMs x 8)."""
import os, sys
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network
threads = max(1, len(os.sched_getaffinity(0)) - 2)
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=1184)
for cap in (1024, 1184, 1024, 1184):
    net.start_service(max_batch=cap, max_wait_us=200)
    net.selfplay(threads=threads, total_moves=512, simulations=160, leaf_batch=16, mate_depth=7, seed=99)
    st = net.selfplay(threads=threads, total_moves=24000, simulations=800, leaf_batch=16, mate_depth=7, seed=1, games_per_thread=12)
    net.stop_service()
    print(f"cap {cap}: {st['moves_per_s']:.0f} moves/s, {st['evals_per_s']/1e6:.2f} M evals/s, mean batch {st['mean_gpu_batch']:.0f}", flush=True)
