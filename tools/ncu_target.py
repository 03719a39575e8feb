"""Short ncu target: 3 warm-up + 5 forwards of batch 1024 (2 launches each: pack_input, tower)."""
import sys
import torch
sys.path.insert(0, ".")
from erweitern_55_b200 import synth, weights as W
from erweitern_55_b200.network import Network

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = Network(precision="bf16", weights=W.init_weights(seed=55), max_batch=n)
x = torch.from_numpy(synth.synthetic_planes(n, seed=1000)).cuda()
p = torch.empty((n, 1725), device="cuda"); v = torch.empty((n, 1), device="cuda")
for _ in range(8):
    net.predict_device(x, p, v)
net.sync()
print("done", float(p.abs().sum()))
