"""GPU bring-up of the tcgen05 tower kernel: layer-wise activations vs the fp32 path, then full outputs."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network
from oracle import network_oracle as O

pert = "--pert" in sys.argv
w = W.init_weights(perturbed=pert)
ref = Network(precision="fp32", weights=w)
print("fp32 net ok", flush=True)
net = Network(precision="fp32", weights=w)
net.set_precision("bf16")
for n in (1, 4, 5, 8, 13, 40):
    x = synth.synthetic_planes(n, seed=1000 + n)
    for nc in (1, 2, 3, 15):
        a = net.debug_activations(x, nc)
        b = ref.debug_activations(x, nc)
        print(f"n={n} convs={nc} max|d|={abs(a-b).max():.4f} ref max={abs(b).max():.3f} nan={np.isnan(a).sum()}", flush=True)
    p, v = net.predict(x)
    po, vo = O.forward_torch(w, x)
    pb, vb = O.forward_bf16_emulated(w, x)
    print(f"n={n} predict: vs fp32 oracle dlogit={abs(p-po).max():.4f} dv={abs(v-vo).max():.5f} | vs bf16 emu dlogit={abs(p-pb).max():.5f} dv={abs(v-vb).max():.6f} top1={(p.argmax(1)==po.argmax(1)).mean():.3f}", flush=True)
import torch
for n in (256, 1024, 4096):
    x = synth.synthetic_planes(n, seed=5)
    p, v = net.predict(x)
    pb, vb = O.forward_bf16_emulated(w, x[:64])
    print(f"n={n} first64 vs bf16 emu dlogit={abs(p[:64]-pb).max():.5f} dv={abs(v[:64]-vb).max():.6f}", flush=True)
    xd = torch.from_numpy(x).cuda()
    pd = torch.empty((n, 1725), device="cuda"); vd = torch.empty((n, 1), device="cuda")
    st = torch.cuda.ExternalStream(net.cuda_stream())
    torch.cuda.synchronize()
    for _ in range(3):
        net.predict_device(xd, pd, vd)
    net.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(20):
            net.predict_device(xd, pd, vd)
        e1.record(st)
    net.sync()
    ms = e0.elapsed_time(e1) / 20
    print(f"n={n} device forward {ms*1e3:.1f} us  {n/ms*1e3:.0f} evals/s  {n/ms*1e3*126368256/1e12:.1f} TFLOP/s algorithmic", flush=True)
    assert abs(pd.cpu().numpy() - p).max() == 0
