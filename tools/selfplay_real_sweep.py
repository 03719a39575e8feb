"""Real self-play (native engine + MCTS) throughput vs host threads x games per thread x service batch."""
import os
import sys
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=2048)
moves = int(os.environ.get("SWEEP_MOVES", "6000"))
configs = [(1024, 200, 128, 1), (1024, 200, 16, 8), (1024, 200, 16, 16), (1024, 200, 32, 8), (512, 150, 16, 8), (512, 150, 16, 16),
           (512, 150, 32, 8), (2048, 300, 16, 16), (2048, 300, 32, 16), (256, 100, 16, 8), (1024, 200, 15, 12), (1024, 200, 14, 12)]
if len(sys.argv) > 1:
    configs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for cap, wait, threads, gpt in configs:
    net.start_service(max_batch=cap, max_wait_us=wait)
    net.selfplay(threads, 512, 160, 16, 7, 1, games_per_thread=gpt)
    st = net.selfplay(threads, moves, 800, 16, 7, 2, games_per_thread=gpt)
    print(f"cap={cap} wait={wait} threads={threads:4d} x games={gpt:3d}: {st['moves_per_s']:8.1f} moves/s {st['evals_per_s']:10.0f} evals/s "
          f"batch {st['mean_gpu_batch']:7.1f} games {st['games']} ({st['seconds']:.2f}s)", flush=True)
    net.stop_service()
