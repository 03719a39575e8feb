"""Device-resident throughput with one, two and three forwards in flight (contexts = streams), float planes and
bit-packed planes: python tools/two_stream_probe.py [batch]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from erweitern_55_b200 import packing, synth, weights as W
from erweitern_55_b200.network import Network
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
w = W.init_weights(seed=55)
nets = [Network(precision="bf16", weights=w, max_batch=B) for _ in range(3)]
host = [synth.synthetic_planes(B, seed=10 + i) for i in range(9)]
xs = [torch.from_numpy(h).cuda() for h in host]
packed = [packing.pack_planes(h) for h in host]
pb = [torch.from_numpy(b.view(np.int32)).cuda() for b, _ in packed]; psc = [torch.from_numpy(s).cuda() for _, s in packed]
ps = [torch.empty((B, 1725), device="cuda") for _ in range(3)]; vs = [torch.empty((B, 1), device="cuda") for _ in range(3)]
def step(i, k, use_packed):
    n = nets[i % k]
    if use_packed:
        n._check(n._lib.e55_predict_packed_device(n._ctx, pb[i % 9].data_ptr(), psc[i % 9].data_ptr(), B, ps[i % k].data_ptr(), vs[i % k].data_ptr()))
    else:
        n.predict_device(xs[i % 9], ps[i % k], vs[i % k])
def run(k, use_packed, steps=3000):
    for i in range(50): step(i, k, use_packed)
    for n in nets: n.sync()
    t0 = time.perf_counter()
    for i in range(steps): step(i, k, use_packed)
    for n in nets: n.sync()
    dt = time.perf_counter() - t0
    print(f"batch {B}, {'packed' if use_packed else 'float '} input, {k} in flight: {dt / steps * 1e6:.1f} us/step = {B * steps / dt / 1e6:.2f} M evals/s", flush=True)
for use_packed in (False, True):
    for k in (1, 2, 3): run(k, use_packed)
