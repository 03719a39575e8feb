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
if len(ib) == 0:          # latency mode: one group per CTA
    ib = ia
t0 = min(ia[0, 0], ib[0, 0])
print("chunks", len(ia), len(ib), "events", len(ep))
print("issuer A: total", ia[-1, 3] - t0, " B:", ib[-1, 3] - t0)
def stats(name, ia):
    cad = np.diff(ia[:, 0])
    print(f"{name}: cadence median={np.median(cad):.0f} mean={cad.mean():.0f} | waitFull med={np.median(ia[:,1]-ia[:,0]):.0f} sum={np.sum(ia[:,1]-ia[:,0])} "
          f"| waitAct med={np.median(ia[:,2]-ia[:,1]):.0f} sum={np.sum(ia[:,2]-ia[:,1])} | issue med={np.median(ia[:,3]-ia[:,2]):.0f} sum={np.sum(ia[:,3]-ia[:,2])} "
          f"| loop med={np.median(ia[1:,0]-ia[:-1,3]):.0f} sum={np.sum(ia[1:,0]-ia[:-1,3])}")
stats("A", ia); stats("B", ib)
big = [(c, int(ia[c,1]-ia[c,0]), int(ia[c,2]-ia[c,1])) for c in range(len(ia)) if ia[c,2]-ia[c,0] > 1000]
print("A stalls >1000 (chunk, waitFull, waitAct):", big[:40])
big = [(c, int(ib[c,1]-ib[c,0]), int(ib[c,2]-ib[c,1])) for c in range(len(ib)) if ib[c,2]-ib[c,0] > 1000]
print("B stalls >1000:", big[:40])
ep = ep[np.argsort(ep[:, 0])]
print("epilogue durations: median", np.median(ep[:,1]-ep[:,0]), "max", (ep[:,1]-ep[:,0]).max())
for c in list(range(0, 24)) + list(range(36, 44)):
    print(f"c={c:3d} A t0={ia[c,0]-t0:7d} wF={ia[c,1]-ia[c,0]:5d} wA={ia[c,2]-ia[c,1]:5d} is={ia[c,3]-ia[c,2]:5d} | B t0={ib[c,0]-t0:7d} wF={ib[c,1]-ib[c,0]:5d} wA={ib[c,2]-ib[c,1]:5d} is={ib[c,3]-ib[c,2]:5d}")
print("layer starts (first chunk of each layer), group A t0 / B t0, and period:")
firsts = [0, 15] + list(range(21, len(ia), 6))
prev = None
for c in firsts:
    if c < len(ia) and c < len(ib):
        print(f"  c={c:3d} A={ia[c,0]-t0:7d} B={ib[c,0]-t0:7d}" + (f"  dA={ia[c,0]-prev[0]:6d} dB={ib[c,0]-prev[1]:6d}" if prev else ""))
        prev = (ia[c,0], ib[c,0])
print("epilogue events (start, duration, group, layer):")
for e in ep[:40]:
    print(f"  t={e[0]-t0:7d} dur={e[1]-e[0]:5d} g={e[2]} L={e[3]}")
cv = a[16 + 8192 + 512:16 + 8192 + 512 + 64].reshape(-1, 4)
cv = cv[cv[:, 0] > 0]; cv = cv[np.argsort(cv[:, 0])]
print("input conversions of thread 0: loop top, plane half free (+wait), loads+stores done (+), fence+arrive done (+)")
for e in cv:
    print(f"  top={e[2]-t0:7d} free={e[0]-t0:7d} (+{e[0]-e[2]:5d}) stored={e[3]-t0:7d} (+{e[3]-e[0]:5d}) arrived={e[1]-t0:7d} (+{e[1]-e[3]:5d})")
