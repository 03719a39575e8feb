"""Bring-up of the 256-filter (scaled tower) tcgen05 path: parity on a 2-block net, timing of 20x256."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network
from oracle import network_oracle as O

w = W.init_weights(blocks=2, filters=256, seed=7, perturbed=True)
ref = Network(blocks=2, filters=256, precision="fp32", weights=w)
net = Network(blocks=2, filters=256, precision="bf16", weights=w)
for n in (1, 4, 5, 13, 40):
    x = synth.synthetic_planes(n, seed=1000 + n)
    for nc in (1, 2, 3, 5):
        a = net.debug_activations(x, nc); b = ref.debug_activations(x, nc)
        print(f"n={n} convs={nc} max|d|={abs(a-b).max():.4f} ref max={abs(b).max():.3f} nan={np.isnan(a).sum()}", flush=True)
    p, v = net.predict(x)
    po, vo = O.forward_torch(w, x)
    print(f"n={n} predict: dlogit={abs(p-po).max():.4f} dv={abs(v-vo).max():.5f} top1={(p.argmax(1)==po.argmax(1)).mean():.3f} logit range {po.min():.2f}..{po.max():.2f}", flush=True)
ref.close(); net.close()

FLOP = 1_240_685_056
w = W.init_weights(blocks=20, filters=256, seed=55)
net = Network(blocks=20, filters=256, precision="bf16", weights=w, max_batch=4096)
x16 = synth.synthetic_planes(16, seed=3)
p, v = net.predict(x16)
po, vo = O.forward_torch(w, x16)
print(f"20x256 n=16: dlogit={abs(p-po).max():.4f} dv={abs(v-vo).max():.5f} top1={(p.argmax(1)==po.argmax(1)).mean():.3f}", flush=True)
st = torch.cuda.ExternalStream(net.cuda_stream())
for n in (16, 592, 1184, 4096):
    x = torch.from_numpy(synth.synthetic_planes(n, seed=5)).cuda()
    pd = torch.empty((n, 1725), device="cuda"); vd = torch.empty((n, 1), device="cuda")
    for _ in range(3): net.predict_device(x, pd, vd)
    net.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record(st)
    for _ in range(reps): net.predict_device(x, pd, vd)
    e1.record(st); net.sync()
    ms = e0.elapsed_time(e1) / reps
    print(f"20x256 n={n}: {ms*1e3:.1f} us  {n/ms*1e3:.0f} evals/s  {n/ms*1e3*FLOP/1e12:.1f} TFLOP/s algorithmic", flush=True)
