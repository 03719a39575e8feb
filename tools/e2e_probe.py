"""Host-buffer predict throughput at batch 1024 (the bench's e2e) and the PCIe copy rates that bound it."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from erweitern_55_b200 import synth, weights as W, packing
from erweitern_55_b200.network import Network
B = 1024
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=B)
x = torch.from_numpy(synth.synthetic_planes(B, seed=1)).pin_memory()
pol = torch.empty((B, 1725), dtype=torch.float32).pin_memory(); val = torch.empty((B, 1), dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda"); dp = torch.empty((B, 1725), device="cuda")
def t(fn, n=30):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
h2d = t(lambda: d.copy_(x, non_blocking=True)); d2h = t(lambda: pol.copy_(dp, non_blocking=True))
print(f"H2D {x.numel()*4/h2d/1e9:.1f} GB/s ({h2d*1e6:.0f} us for {x.numel()*4/1e6:.1f} MB); D2H {pol.numel()*4/d2h/1e9:.1f} GB/s ({d2h*1e6:.0f} us)")
lib, ctx = net._lib, net._ctx
e = t(lambda: lib.e55_predict(ctx, x.data_ptr(), B, pol.data_ptr(), val.data_ptr()), 50)
print(f"e55_predict (fp32 planes): {e*1e6:.0f} us/step = {B/e/1e6:.2f} M evals/s")
bits, scale = packing.pack_planes(x.numpy())
bt = torch.from_numpy(bits.astype(np.int32)).pin_memory(); st = torch.from_numpy(scale).pin_memory()
e = t(lambda: lib.e55_predict_packed(ctx, bt.data_ptr(), st.data_ptr(), B, pol.data_ptr(), val.data_ptr()), 50)
print(f"e55_predict_packed: {e*1e6:.0f} us/step = {B/e/1e6:.2f} M evals/s")
for threads in (0, 4, 8, 12, 14, 16):
    lib.e55_set_host_threads(ctx, threads)
    e = t(lambda: lib.e55_predict(ctx, x.data_ptr(), B, pol.data_ptr(), val.data_ptr()), 50)
    print(f"e55_predict, {threads:2d} packing threads: {e*1e6:.0f} us/step = {B/e/1e6:.2f} M evals/s")
