"""Host-buffer predict at batch 1024 over rotating input buffers (as bench.py's e2e): every fixed split between packed
and float chunks (E55_HOST_SPLIT), then the library's automatic choice."""
import ctypes as C, os, sys, time
sys.path.insert(0, ".")
import torch
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network
B = 1024
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=B)
xs = [torch.from_numpy(synth.synthetic_planes(B, seed=i)).pin_memory() for i in range(3)]
pol = torch.empty((B, 1725), dtype=torch.float32).pin_memory(); val = torch.empty((B, 1), dtype=torch.float32).pin_memory()
lib, ctx = net._lib, net._ctx
def run(n=60):
    for i in range(14): lib.e55_predict(ctx, xs[i % 3].data_ptr(), B, pol.data_ptr(), val.data_ptr())
    t0 = time.perf_counter()
    for i in range(n): lib.e55_predict(ctx, xs[i % 3].data_ptr(), B, pol.data_ptr(), val.data_ptr())
    return (time.perf_counter() - t0) / n
for rep in range(2):
    for split in (0, 1, 2, 3, 4, 8):
        os.environ["E55_HOST_SPLIT"] = str(split)
        lib.e55_set_host_threads(ctx, -1)
        e = run()
        print(f"split {split}/8 as floats: {e*1e6:.0f} us/step = {B/e/1e6:.2f} M evals/s", flush=True)
del os.environ["E55_HOST_SPLIT"]
lib.e55_set_host_threads(ctx, -1)
e = run(200)
us = (C.c_double * 6)(); eighths = C.c_int()
lib.e55_get_host_split(ctx, C.byref(eighths), us)
print(f"automatic: {e*1e6:.0f} us/step = {B/e/1e6:.2f} M evals/s; last share {eighths.value}/8; us/position {[round(u, 3) for u in us]}")
