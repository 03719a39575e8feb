"""Latency / throughput sweep over batch sizes (BASELINE.json configs[2]): device-resident forward
(CUDA events on the context stream) and the host-buffer Network.predict call (wall clock, numpy in/out)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network

net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=4096)
st = torch.cuda.ExternalStream(net.cuda_stream())
rows = []
for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 1184, 2048, 4096):
    xh = synth.synthetic_planes(n, seed=1000 + n)
    x = torch.from_numpy(xh).cuda()
    p = torch.empty((n, 1725), device="cuda"); v = torch.empty((n, 1), device="cuda")
    for _ in range(5):
        net.predict_device(x, p, v)
    net.sync()
    reps = 200 if n <= 1184 else 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        net.predict_device(x, p, v)
    e1.record(st)
    net.sync()
    dev_us = e0.elapsed_time(e1) / reps * 1e3
    net.predict(xh)
    hreps = 50 if n <= 256 else 10
    t0 = time.perf_counter()
    for _ in range(hreps):
        net.predict(xh)
    host_us = (time.perf_counter() - t0) / hreps * 1e6
    rows.append({"batch": n, "device_us": dev_us, "device_evals_per_s": n / dev_us * 1e6,
                 "host_predict_us": host_us, "host_evals_per_s": n / host_us * 1e6})
    print(f"batch {n:5d}: device {dev_us:8.1f} us {n/dev_us*1e6:12.0f} evals/s | Network.predict (numpy, pageable) {host_us:9.1f} us {n/host_us*1e6:11.0f} evals/s", flush=True)
json.dump(rows, open("gpurun_out/batch_sweep.json", "w"), indent=1)
