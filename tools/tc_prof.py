"""In-kernel cycle counters of the tower kernel (CTA 0) for one forward of n positions."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network

w = W.init_weights()
net = Network(precision="bf16", weights=w)
import os
PACKED = os.environ.get("PACKED") == "1"      # device-resident bit-packed input (what the self-play paths and the host pipeline feed)
from erweitern_55_b200 import packing
lag = int(os.environ.get("E55_LAG", "-1"))
if lag >= 0:
    net._check(net._lib.e55_debug_set_lag(net._ctx, lag))
for n in [int(a) for a in sys.argv[1:]] or [8, 256, 1024]:
    xh = synth.synthetic_planes(n, seed=5)
    x = torch.from_numpy(xh).cuda()
    pd = torch.empty((n, 1725), device="cuda"); vd = torch.empty((n, 1), device="cuda")
    if PACKED:
        bits, scale = packing.pack_planes(xh)
        db, ds = torch.from_numpy(bits.view(np.int32)).cuda(), torch.from_numpy(scale).cuda()
        fwd = lambda: net._check(net._lib.e55_predict_packed_device(net._ctx, db.data_ptr(), ds.data_ptr(), n, pd.data_ptr(), vd.data_ptr()))
        net.predict_device = lambda *_: fwd()
    net.predict_device(x, pd, vd); net.sync()
    net._check(net._lib.e55_debug_profile(net._ctx, 1, None))
    net.predict_device(x, pd, vd); net.sync()
    out = (C.c_uint64 * (16 + 8192 + 1024))()
    net._check(net._lib.e55_debug_profile(net._ctx, 0, out))
    c = list(out)
    print(f"n={n}: issuerA total={c[0]} issue={c[1]} wait_full={c[2]} wait_act={c[3]} | issuerB total={c[4]} issue={c[5]} "
          f"wait_full={c[6]} wait_act={c[7]} | producer wait={c[8]} | epi wait={c[10]} work={c[11]}", flush=True)
    print(f"   kernel timeline (cycles from entry): setup done={c[13]-c[12]} issuerA start={c[9]-c[12]} issuerA end={c[9]+c[0]-c[12]} "
          f"all roles done={c[14]-c[12]} exit={c[15]-c[12]}", flush=True)
    st = torch.cuda.ExternalStream(net.cuda_stream())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        net.predict_device(x, pd, vd)
    net.sync()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(20):
            net.predict_device(x, pd, vd)
        e1.record(st)
    net.sync()
    ms = e0.elapsed_time(e1) / 20
    print(f"   lag={lag} n={n}: {ms*1e3:.1f} us/forward  {n/ms*1e3:.0f} evals/s", flush=True)
