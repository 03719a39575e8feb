"""Per-layer activations of the channel-split latency kernel against the fp32 CUDA path: first wrong layer, which channels."""
import sys
sys.path.insert(0, ".")
import numpy as np
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network
w = W.init_weights(seed=55, perturbed=True)
a = Network(precision="bf16", weights=w); b = Network(precision="fp32", weights=w)
x = synth.synthetic_planes(5, seed=1005)
import os
if os.environ.get("DUMP_RANK"):
    a._check(a._lib.e55_debug_set_lag(a._ctx, int(os.environ["DUMP_RANK"])))
for nconv in (1, 2, 3, 4, 15):
    ya, yb = a.debug_activations(x, nconv), b.debug_activations(x, nconv)
    d = np.abs(ya - yb)
    print(f"after {nconv} convs: max diff {d.max():.4f} (|ref| max {np.abs(yb).max():.2f}); per 32-channel block: "
          f"{[round(float(d[:, k*32:(k+1)*32].max()), 3) for k in range(4)]}; per position {np.round(d.max(axis=(1,2,3)), 3)}")
ya, yb = a.debug_activations(x, 1), b.debug_activations(x, 1)
np.set_printoptions(precision=3, suppress=True, linewidth=200)
print("split ch64 pos0:", ya[0, 64].ravel()[:12]); print("ref   ch64 pos0:", yb[0, 64].ravel()[:12]); print("ref   ch0  pos0:", yb[0, 0].ravel()[:12])
print("split ch96 pos0:", ya[0, 96].ravel()[:12]); print("ref   ch96 pos0:", yb[0, 96].ravel()[:12]); print("ref   ch32 pos0:", yb[0, 32].ravel()[:12])
for c in (64, 65, 96, 127):
    best = min(range(128), key=lambda k: float(np.abs(ya[:, c] - yb[:, k]).max()))
    print("split channel", c, "is closest to ref channel", best, "max diff", float(np.abs(ya[:, c] - yb[:, best]).max()))
