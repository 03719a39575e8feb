"""What the evaluation service itself sustains (no tree work): asynchronous 16-position requests from N host threads."""
import sys
sys.path.insert(0, ".")
from erweitern_55_b200 import weights as W
from erweitern_55_b200.network import Network
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=2048)
for cap, wait in ((512, 150), (1024, 200), (2048, 300)):
    net.start_service(max_batch=cap, max_wait_us=wait)
    for workers in (2, 4, 8, 14):
        net.bench_selfplay(workers, 2, 160, 16)
        mv, ev, mb = net.bench_selfplay(workers, 200, 800, 16)
        print(f"cap={cap} workers={workers:3d}: {ev:10.0f} evals/s  mean GPU batch {mb:7.1f}", flush=True)
    net.stop_service()
