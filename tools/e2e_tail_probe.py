"""A/B of the host pipeline's tail schedule: Network.predict at batch 1024, pageable input (all packed)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network
B = 1024
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=B)
xs = [synth.synthetic_planes(B, seed=10 + i) for i in range(3)]
for i in range(30): net.predict(xs[i % 3])
best = []
for rep in range(5):
    t0 = time.perf_counter()
    for i in range(100): net.predict(xs[i % 3])
    best.append((time.perf_counter() - t0) / 100)
print("us/call", [round(b * 1e6) for b in best], "M evals/s", round(B / min(best) / 1e6, 3))
