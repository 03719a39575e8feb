"""The drop-in call itself -- Network.predict(numpy) -> numpy -- at the reference's batch sizes and at batch 1024, with
pageable and with page-locked input arrays, plus the raw copy rates that bound it."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from erweitern_55_b200 import hostmem, synth, weights as W
from erweitern_55_b200.network import Network

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=B)
lib, ctx = net._lib, net._ctx
print("cores", len(os.sched_getaffinity(0)))


def t(fn, n):
    for i in range(3):
        fn(i)
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    return (time.perf_counter() - t0) / n


for nb in (1, 16, 64, 256, 1024, 1184):
    xs = [synth.synthetic_planes(nb, seed=10 + i) for i in range(3)]
    xp = []
    for a in xs:
        b = hostmem.empty(a.shape); b[...] = a; xp.append(b)
    reps = 300 if nb <= 64 else 40
    for name, arrs in (("pageable", xs), ("pinned", xp)):
        lib.e55_set_host_threads(ctx, -1)
        for i in range(20):
            net.predict(arrs[i % 3])
        c = t(lambda i: net.predict(arrs[i % 3]), reps)
        extra = ""
        if nb >= 512 and name == "pinned":
            us = (C.c_double * 8)(); e8 = C.c_int()
            lib.e55_get_host_split(ctx, C.byref(e8), us)
            extra = " split us/pos " + " ".join(f"{s}:{u:.3f}" for s, u in zip((0, 1, 2, 3, 4, 5, 6, 8), us))
        print(f"Network.predict batch {nb:5d} {name:8s} input: {c * 1e6:7.0f} us/call = {nb / c / 1e6:.3f} M evals/s{extra}", flush=True)
    if nb >= 256:
        for threads in (0, 3, 7, 15):
            lib.e55_set_host_threads(ctx, threads)
            c = t(lambda i: net.predict(xs[i % 3]), reps)
            print(f"   pageable, {threads:2d} pool threads: {c * 1e6:7.0f} us/call = {nb / c / 1e6:.3f} M evals/s", flush=True)
net.close()
