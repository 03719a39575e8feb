"""Real self-play (native engine + MCTS) throughput vs game threads / service batch."""
import sys
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=2048)
for cap, wait in ((1024, 200), (2048, 300)):
    net.start_service(max_batch=cap, max_wait_us=wait)
    for threads in (64, 128, 256, 512):
        net.selfplay(threads, threads, 160, 16, 7, 1)
        st = net.selfplay(threads, threads * 8, 800, 16, 7, 2)
        print(f"cap={cap} threads={threads:4d}: {st['moves_per_s']:8.1f} moves/s {st['evals_per_s']:10.0f} evals/s batch {st['mean_gpu_batch']:7.1f} games {st['games']} ({st['seconds']:.2f}s)", flush=True)
    net.stop_service()
