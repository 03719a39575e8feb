"""Per-layer timeline of the channel-split latency kernel (CTA 0): epilogue thread 0 and the MMA issuer."""
import ctypes as C, os, sys
os.environ["E55_SPLIT_PROF"] = "1"
import numpy as np, torch
sys.path.insert(0, ".")
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net = Network(precision="bf16", weights=W.init_weights())
x = torch.from_numpy(synth.synthetic_planes(n, seed=5)).cuda()
pd = torch.empty((n, 1725), device="cuda"); vd = torch.empty((n, 1), device="cuda")
net.predict_device(x, pd, vd); net.sync()
net._check(net._lib.e55_debug_profile(net._ctx, 1, None))
net.predict_device(x, pd, vd); net.sync()
out = (C.c_uint64 * (16 + 8192 + 1024))()
net._check(net._lib.e55_debug_profile(net._ctx, 0, out))
a = np.array(list(out), dtype=np.int64)[16:16 + 17 * 8].reshape(17, 8)
t0 = a[0, 5]
print("layer | issuer: first chunk reached, waits over | epilogue: waiting since, accumulator complete (+wait), math done, peer done (+wait), stores + copies issued")
for L in range(17):
    r = a[L]
    print(f"{L:3d}   | {r[5]-t0:7d} {r[6]-t0:7d} (+{r[6]-r[5]:5d}) | {r[0]-t0:7d} {r[1]-t0:7d} (+{r[1]-r[0]:5d}) {r[2]-t0:7d} {r[3]-t0:7d} (+{r[3]-r[2]:5d}) {r[4]-t0:7d}")
