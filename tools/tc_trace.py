"""Timeline of CTA 0's first round: per-chunk issuer timestamps and per-event epilogue timestamps."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = Network(precision="bf16", weights=W.init_weights())
lag = int(os.environ.get("E55_LAG", "-1"))
if lag >= 0:
    net._check(net._lib.e55_debug_set_lag(net._ctx, lag))
x = torch.from_numpy(synth.synthetic_planes(n, seed=5)).cuda()
pd = torch.empty((n, 1725), device="cuda"); vd = torch.empty((n, 1), device="cuda")
net.predict_device(x, pd, vd); net.sync()
net._check(net._lib.e55_debug_profile(net._ctx, 1, None))
net.predict_device(x, pd, vd); net.sync()
NW = 16 + 8192 + 1024
out = (C.c_uint64 * NW)()
net._check(net._lib.e55_debug_profile(net._ctx, 0, out))
a = np.array(list(out), dtype=np.int64)
np.save("gpurun_out/trace.npy", a)
ia = a[16:16 + 4096].reshape(-1, 4); ib = a[16 + 4096:16 + 8192].reshape(-1, 4); ep = a[16 + 8192:].reshape(-1, 4)
ia = ia[ia[:, 0] > 0]; ib = ib[ib[:, 0] > 0]; ep = ep[ep[:, 0] > 0]
t0 = min(ia[0, 0], ib[0, 0])
print("chunks", len(ia), len(ib), "events", len(ep))
print("issuer A: total", ia[-1, 3] - t0, " B:", ib[-1, 3] - t0)
def fmt(r): return " ".join(f"{v - t0:7d}" for v in r)
nch = len(ia)
for c in list(range(0, 40)) + list(range(nch - 16, nch)):
    print(f"c={c:3d} A[{fmt(ia[c])}] dWaitFull={ia[c,1]-ia[c,0]:5d} dWaitAct={ia[c,2]-ia[c,1]:5d} dIssue={ia[c,3]-ia[c,2]:5d} | B[{fmt(ib[c])}] dWF={ib[c,1]-ib[c,0]:5d} dWA={ib[c,2]-ib[c,1]:5d} dI={ib[c,3]-ib[c,2]:5d}")
ep = ep[np.argsort(ep[:, 0])]
for e in ep[:24]:
    print(f"epi g={e[2]} L={e[3]:2d} ready={e[0]-t0:7d} done={e[1]-t0:7d} dur={e[1]-e[0]}")
