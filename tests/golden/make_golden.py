"""Generates tests/golden/forward_golden.npz from the CPU oracle (run here, committed with its output).

The reference's own forward pass (TensorFlow 1.15.2) cannot run in this environment and the reference
holds no golden vectors for it, so these vectors pin the ORACLE and the weight /
input generators against silent drift; they are not outputs of the reference.  fp64 twin outputs are
stored as float64; the bf16-emulated outputs as float32.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from erweitern_55_b200 import synth, weights as W  # noqa: E402
from oracle import network_oracle as O  # noqa: E402

INPUT_SEED = 1234
BATCH = 4


def build():
    out = {"input_seed": np.int64(INPUT_SEED), "batch": np.int64(BATCH)}
    x = synth.synthetic_planes(BATCH, seed=INPUT_SEED)
    out["input_checksum"] = np.float64(x.astype(np.float64).sum())
    for tag, pert in (("default", False), ("perturbed", True)):
        w = W.init_weights(seed=55, perturbed=pert)
        out[f"{tag}_weight_sums"] = np.array([t.astype(np.float64).sum() for t in w])
        p64, v64 = O.forward_torch(w, x, torch.float64)
        pb, vb = O.forward_bf16_emulated(w, x)
        out[f"{tag}_policy_f64"], out[f"{tag}_value_f64"] = p64, v64
        out[f"{tag}_policy_bf16"], out[f"{tag}_value_bf16"] = pb, vb
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "forward_golden.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
