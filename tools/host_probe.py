"""Facts about the host side of a GPU box that bound the host-buffer predict: cores, topology, copy rates, and the
drop-in call itself (Network.predict on pageable numpy arrays) at the reference's batch sizes and at batch 1024."""
import os
import subprocess
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as exc:
        return f"<{exc}>"


print(sh("lscpu | egrep 'Model name|Socket|Core|Thread|^CPU\\(s\\)|NUMA|L2|L3|Flags' | cut -c1-400"))
print("affinity:", len(os.sched_getaffinity(0)), "cpus")
print(sh("nvidia-smi topo -m | head -20"))
print(sh("free -g | head -2"))

B = 1024
x = synth.synthetic_planes(B, seed=1)
xs = [np.roll(x, i, axis=0).copy() for i in range(3)]   # 81 MB: not cache resident


def t(fn, n=20):
    fn(0); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n


dst = np.empty_like(x)
c = t(lambda i: np.copyto(dst, xs[i % 3]))
print(f"numpy memcpy 27 MB, one thread: {x.nbytes / c / 1e9:.1f} GB/s")
xp = [torch.from_numpy(a).pin_memory() for a in xs]
d = torch.empty((B, 266, 5, 5), device="cuda")
c = t(lambda i: d.copy_(xp[i % 3], non_blocking=True))
print(f"H2D pinned: {x.nbytes / c / 1e9:.1f} GB/s ({c * 1e6:.0f} us)")
xt = [torch.from_numpy(a) for a in xs]
c = t(lambda i: d.copy_(xt[i % 3]))
print(f"H2D pageable: {x.nbytes / c / 1e9:.1f} GB/s ({c * 1e6:.0f} us)")
dp = torch.empty((B, 1725), device="cuda")
pp = torch.empty((B, 1725)).pin_memory()
c = t(lambda i: pp.copy_(dp, non_blocking=True))
print(f"D2H pinned 7 MB: {pp.numel() * 4 / c / 1e9:.1f} GB/s ({c * 1e6:.0f} us)")
c = t(lambda i: np.empty((B, 1725), dtype=np.float32).fill(0))
print(f"np.empty 7 MB + first touch: {c * 1e6:.0f} us")

net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=B)
for nb in (1, 16, 64, 256, 1024, 1184):
    xb = [synth.synthetic_planes(nb, seed=10 + i) for i in range(3)]
    reps = 200 if nb <= 64 else 30
    for i in range(14):
        net.predict(xb[i % 3])
    c = t(lambda i: net.predict(xb[i % 3]), reps)
    print(f"Network.predict(numpy pageable) batch {nb}: {c * 1e6:.0f} us/call = {nb / c / 1e6:.3f} M evals/s", flush=True)
net.close()
