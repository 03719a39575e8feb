"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), one kernel family per mode:

    compute-sanitizer --tool racecheck python tools/sanitize_target.py tower128

tower128 / tower256: a few forwards of the tcgen05 tower kernel at batch sizes that exercise one group per CTA, both
groups, a ragged last round and more than one round per CTA pair; results are checked against the fp32 CUDA path.
legal: the fused legal-move tail.  selfplay: GPU-resident self-play (select / expand kernels) on a few games.
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network

mode = sys.argv[1] if len(sys.argv) > 1 else "tower128"

if mode in ("tower128", "tower256"):
    blocks, filters = (7, 128) if mode == "tower128" else (2, 256)
    w = W.init_weights(blocks=blocks, filters=filters, seed=55)
    net = Network(precision="bf16", blocks=blocks, filters=filters, weights=w, max_batch=64)
    ref = Network(precision="fp32", blocks=blocks, filters=filters, weights=w, max_batch=64)
    for n in ((1, 5, 16, 37) if mode == "tower128" else (1, 9)):
        x = synth.synthetic_planes(n, seed=1000 + n)
        p, v = net.predict(x)
        pr, vr = ref.predict(x)
        print(mode, "batch", n, "max |dlogit|", float(np.abs(p - pr).max()), "max |dvalue|", float(np.abs(v - vr).max()), flush=True)
        assert np.abs(p - pr).max() < 0.1 and np.abs(v - vr).max() < 0.02
    net.close(); ref.close()
elif mode == "legal":
    from erweitern_55_b200 import packing
    w = W.init_weights(seed=55)
    net = Network(precision="bf16", weights=w, max_batch=64)
    rng = np.random.default_rng(3)
    for n in (1, 16, 23):
        x = synth.synthetic_planes(n, seed=2000 + n)
        bits, scale = packing.pack_planes(x)
        counts = rng.integers(0, 60, size=n)
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        idx = np.concatenate([rng.choice(1725, size=c, replace=False) for c in counts] + [np.zeros(0, dtype=np.int64)]).astype(np.uint16)
        pri, v = net.predict_packed_moves(bits, scale, off, idx)
        p, v2 = net.predict_packed(bits, scale)
        for i in range(n):
            lo, hi = off[i], off[i + 1]
            if hi > lo:
                z = p[i, idx[lo:hi]].astype(np.float64)
                e = np.exp(z - z.max())
                assert np.abs(pri[lo:hi] - e / e.sum()).max() < 1e-5
        print("legal batch", n, "ok", flush=True)
    net.close()
elif mode == "selfplay":
    w = W.init_weights(seed=55)
    net = Network(precision="bf16", weights=w)
    games = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    stats, recs = net.gpu_selfplay(games=games, total_moves=games * 3, simulations=32, records=8, seed=3)
    print("selfplay", stats, len(recs), flush=True)
    net.close()
else:
    raise SystemExit("unknown mode " + mode)
print("done", mode)
