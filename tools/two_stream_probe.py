"""Device-resident throughput at batch 1024 with one and with two forwards in flight (two contexts = two streams)."""
import sys, time
sys.path.insert(0, ".")
import torch
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
w = W.init_weights(seed=55)
nets = [Network(precision="bf16", weights=w, max_batch=B) for _ in range(3)]
xs = [torch.from_numpy(synth.synthetic_planes(B, seed=10 + i)).cuda() for i in range(9)]
ps = [torch.empty((B, 1725), device="cuda") for _ in range(3)]; vs = [torch.empty((B, 1), device="cuda") for _ in range(3)]
def run(k, steps=3000):
    for i in range(50): nets[i % k].predict_device(xs[i % 9], ps[i % k], vs[i % k])
    for n in nets: n.sync()
    t0 = time.perf_counter()
    for i in range(steps): nets[i % k].predict_device(xs[i % 9], ps[i % k], vs[i % k])
    for n in nets: n.sync()
    dt = time.perf_counter() - t0
    print(f"batch {B}, {k} in flight: {dt / steps * 1e6:.1f} us/step = {B * steps / dt / 1e6:.2f} M evals/s", flush=True)
for k in (1, 2, 3, 1, 2): run(k)
