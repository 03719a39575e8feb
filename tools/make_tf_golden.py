#!/usr/bin/env python
"""Reference-side golden vectors for the forward pass -- to be run where the REFERENCE runs (Python 3.6/3.7 with
tensorflow==1.15.2 / tensorflow-gpu==1.15.2 as pinned in the reference's requirements.txt:32; it cannot run in this
repo's containers, which is why the forward-pass oracle is marked "parity unpinned").

    cd <checkout of nyashiki/erweitern_55>/erweitern_55
    python <this repo>/tools/make_tf_golden.py --out <this repo>/tests/golden/tf_golden.npz [--device cpu]

What it does: builds the reference's own ``network.Network`` (network.py:16-73), installs seeded weights through the
reference's ``set_weights`` (so the file also pins the get_weights()/set_weights() tensor order, network.py:228-243),
runs the reference's ``predict`` (network.py:201-216) on seeded synthetic inputs, and stores weights, inputs and outputs.
``tests/test_gpu_parity.py::test_tf_golden_when_present`` and ``tests/test_oracle.py::test_tf_golden_pins_the_oracle``
consume the file: the oracle must reproduce the outputs to 1e-4 / 1e-5 and both CUDA paths are then held to the
reference's outputs themselves.  Everything below is numpy except the three calls into the reference.
"""
import argparse

import numpy as np


def synthetic_planes(n, seed, in_planes=266):
    """Same generator as erweitern_55_b200.synth.synthetic_planes (kept in step by tests/test_host_logic.py)."""
    rng = np.random.default_rng(seed)
    x = (rng.random((n, in_planes, 5, 5)) < 0.1).astype(np.float32)
    x[:, 0] = rng.integers(0, 2, size=(n, 1, 1)).astype(np.float32)
    x[:, 1] = (rng.integers(0, 512, size=(n, 1, 1)) / 512.0).astype(np.float32)
    return x


def perturb(weights, seed):
    """Random BatchNorm statistics, biases and kernels of the reference's own shapes: every tensor of model.get_weights()
    is replaced by seeded values of the kind its shape implies (kernels keep their Glorot scale), so that BN folding, the
    bias paths and the head ordering are all exercised."""
    rng = np.random.default_rng(seed)
    out, i = [], 0
    while i < len(weights):
        w = weights[i]
        if w.ndim in (2, 4):                                       # kernel, then its bias
            fan_in = int(np.prod(w.shape[:-1]))
            fan_out = int(w.shape[-1]) * (int(np.prod(w.shape[:-2])) if w.ndim == 4 else 1)
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            out.append(rng.uniform(-lim, lim, size=w.shape).astype(np.float32))
            out.append(rng.normal(0.0, 0.05, size=weights[i + 1].shape).astype(np.float32))
            i += 2
        else:                                                      # BatchNorm: gamma, beta, moving_mean, moving_variance
            c = w.shape
            out.append(rng.uniform(0.5, 1.5, size=c).astype(np.float32))
            out.append((rng.normal(0.0, 0.1, size=c) if c != (1,) else np.full(c, 0.3)).astype(np.float32))
            out.append(rng.normal(0.0, 0.1, size=c).astype(np.float32))
            out.append(rng.uniform(0.5, 1.5, size=c).astype(np.float32))
            i += 4
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tf_golden.npz")
    ap.add_argument("--device", default="cpu", choices=["cpu", "gpu"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seed", type=int, default=55)
    args = ap.parse_args()
    import network                                                 # the reference's erweitern_55/network.py
    nn = network.Network(args.device)
    weights = perturb(nn.get_weights(), args.seed)
    nn.set_weights(weights)
    back = nn.get_weights()
    assert all(np.array_equal(a, b) for a, b in zip(weights, back)), "set_weights/get_weights do not round-trip"
    x = synthetic_planes(args.batch, seed=1000 + args.batch)
    policy, value = nn.predict(x)
    assert policy.shape == (args.batch, 1725) and value.shape == (args.batch, 1)
    np.savez_compressed(args.out, n_weights=np.int64(len(weights)), inputs=x, policy=np.asarray(policy, dtype=np.float32),
                        value=np.asarray(value, dtype=np.float32), seed=np.int64(args.seed),
                        shapes=np.array([str(tuple(w.shape)) for w in weights]),
                        **{"w%d" % i: w for i, w in enumerate(weights)})
    print("wrote", args.out, "|logit| max", float(np.abs(policy).max()), "value range", float(value.min()), float(value.max()))


if __name__ == "__main__":
    main()
