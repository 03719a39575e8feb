"""Network.predict at the reference's own batch sizes (1 and 16), pageable numpy in and out: us per call."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network
net = Network(precision="bf16", weights=W.init_weights(seed=55))
for nb in (1, 16, 32):
    xs = [synth.synthetic_planes(nb, seed=70 + i) for i in range(3)]
    ref = [net.predict(x) for x in xs]
    for i in range(50): net.predict(xs[i % 3])
    best = []
    for rep in range(5):
        t0 = time.perf_counter()
        for i in range(300): p, v = net.predict(xs[i % 3])
        best.append((time.perf_counter() - t0) / 300 * 1e6)
    ok = all(np.array_equal(net.predict(x)[0], r[0]) and np.array_equal(net.predict(x)[1], r[1]) for x, r in zip(xs, ref))
    print(f"batch {nb}: {min(best):.1f} us/call (runs {[round(b, 1) for b in best]}), results stable: {ok}")
