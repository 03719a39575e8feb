"""Where do the channel-split latency kernel and the throughput kernel differ?  (per policy channel / square / value)"""
import os, sys, subprocess
sys.path.insert(0, ".")
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from erweitern_55_b200 import synth, weights as W
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=W.init_weights(seed=55, perturbed=True))
    out = {}
    for n in (1, 4, 5, 16):
        x = synth.synthetic_planes(n, seed=1000 + n)
        p, v = net.predict(x)
        out[f"p{n}"], out[f"v{n}"] = p, v
    np.savez(sys.argv[2], **out)
    sys.exit(0)
for tag, smax in (("split", "64"), ("pair", "0")):
    subprocess.run([sys.executable, __file__, "child", f"/tmp/diag_{tag}.npz"], env=dict(os.environ, E55_SPLIT_MAX=smax), check=True)
a, b = np.load("/tmp/diag_split.npz"), np.load("/tmp/diag_pair.npz")
for n in (1, 4, 5, 16):
    d = np.abs(a[f"p{n}"] - b[f"p{n}"]).reshape(n, 69, 25)
    print(f"n={n}: max |dlogit| {d.max():.4f}  |dvalue| {np.abs(a[f'v{n}'] - b[f'v{n}']).max():.5f}")
    print("   per position:", np.round(d.max(axis=(1, 2)), 3))
    print("   per channel :", np.round(d.max(axis=(0, 2)), 2))
    print("   per square  :", np.round(d.max(axis=(0, 1)), 2))
