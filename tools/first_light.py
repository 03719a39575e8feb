"""GPU bring-up script: UMMA addressing self-test, tensor-pipe micro-bench, fp32 parity vs the oracle."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from erweitern_55_b200 import _capi, synth, weights as W
from erweitern_55_b200.network import Network
from oracle import network_oracle as O

out = {}
lib = _capi.load()
for shift in (0, 1, 5, 6, 7, 8, 13, 42):
    err = C.c_float()
    rc = lib.e55_selftest_umma(0, shift, C.byref(err))
    out[f"selftest_shift_{shift}"] = (rc, err.value)
    print("selftest shift", shift, rc, err.value, flush=True)
for nd in (128, 256):
    tf = C.c_float()
    rc = lib.e55_bench_umma(0, nd, 200, C.byref(tf))
    out[f"umma_bench_n{nd}"] = (rc, tf.value)
    print("umma bench N", nd, rc, tf.value, "TFLOP/s", flush=True)

for pert in (False, True):
    w = W.init_weights(perturbed=pert)
    net = Network(precision="fp32", weights=w)
    for n in (1, 16, 37, 256):
        x = synth.synthetic_planes(n, seed=1000 + n)
        p, v = net.predict(x)
        po, vo = O.forward_torch(w, x)
        print(f"fp32 pert={pert} n={n} dlogit={abs(p-po).max():.3e} dvalue={abs(v-vo).max():.3e} "
              f"top1={(p.argmax(1)==po.argmax(1)).mean():.4f}", flush=True)
        out[f"fp32_{pert}_{n}"] = (float(abs(p - po).max()), float(abs(v - vo).max()))
    x = synth.synthetic_planes(1024)
    net.predict(x)
    t = time.time(); net.predict(x); dt = time.time() - t
    print("fp32 batch 1024 host predict", dt * 1e3, "ms", 1024 / dt, "evals/s", flush=True)
    m = synth.synthetic_legal_mask(16)
    x = synth.synthetic_planes(16)
    pr, v = net.predict_masked(x, m)
    po, vo = O.forward_torch(w, x)
    print("masked softmax err", abs(pr - O.masked_softmax(po, m)).max(), flush=True)
    net.close()
json.dump(out, open("gpurun_out/first_light.json", "w"), indent=1)
