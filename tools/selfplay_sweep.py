"""Synthetic self-play load (tree ops excluded) vs number of worker threads, through the aggregation service."""
import sys
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=2048)
for cap, wait in ((1024, 200), (512, 100), (2048, 300)):
    net.start_service(max_batch=cap, max_wait_us=wait)
    for workers in (32, 64, 128, 256):
        net.bench_selfplay(workers, 1, 160, 16)
        mv, ev, mb = net.bench_selfplay(workers, 3, 800, 16)
        print(f"cap={cap} wait={wait}us workers={workers:4d}: {mv:8.1f} moves/s  {ev:10.0f} evals/s  mean GPU batch {mb:7.1f}", flush=True)
    net.stop_service()
