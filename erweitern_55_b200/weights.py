"""Keras ``get_weights()`` list format of the reference network and synthetic initialisers.

The reference hands weights around as the flat list ``keras.Model.get_weights()`` returns
(reference ``erweitern_55/network.py:228-243``; wire use ``client.py:50-51``, ``train.py:50``).
This module describes that list for the AlphaZero-style tower of ``network.py:75-146`` and builds
synthetic lists with the Keras-1.15 default initialisers (glorot-uniform kernels, zero biases,
BatchNorm gamma=1/beta=0/mean=0/var=1) or a "perturbed" variant that exercises the BN fold and
bias paths.  No arithmetic of the forward pass lives here; folding and packing happen inside the
C-ABI library (``csrc/e55_api.cu``).

Order of tensors (Keras sorts functional-model layers by graph depth, ties by creation order):
    stem conv (kernel HWIO, bias), stem BN (gamma, beta, moving_mean, moving_variance),
    per residual block: conv, BN, conv, BN,
    value conv1x1, value BN, policy conv3x3, policy BN, value dense 25->128,
    policy conv1x1 -> 69, value dense 128->1.
"""
from __future__ import annotations

import numpy as np

IN_PLANES = 266      # network.py:85
BOARD = 5            # 5x5 minishogi board
SQUARES = BOARD * BOARD
POLICY_CHANNELS = 69  # network.py:106-107
POLICY_SIZE = POLICY_CHANNELS * SQUARES  # 1725, network.py:108
VALUE_HIDDEN = 128   # network.py:116-117
BN_EPS = 1e-3        # Keras BatchNormalization default epsilon


def weight_spec(blocks: int = 7, filters: int = 128, in_planes: int = IN_PLANES):
    """Return ``[(name, shape), ...]`` in Keras ``get_weights()`` order."""
    spec = []

    def conv(name, kh, cin, cout):
        spec.append((name + "/kernel", (kh, kh, cin, cout)))
        spec.append((name + "/bias", (cout,)))

    def bn(name, c):
        for part in ("gamma", "beta", "moving_mean", "moving_variance"):
            spec.append((name + "/" + part, (c,)))

    def dense(name, cin, cout):
        spec.append((name + "/kernel", (cin, cout)))
        spec.append((name + "/bias", (cout,)))

    conv("stem_conv", 3, in_planes, filters)
    bn("stem_bn", filters)
    for b in range(blocks):
        conv(f"block{b}_conv1", 3, filters, filters)
        bn(f"block{b}_bn1", filters)
        conv(f"block{b}_conv2", 3, filters, filters)
        bn(f"block{b}_bn2", filters)
    conv("value_conv", 1, filters, 1)
    bn("value_bn", 1)
    conv("policy_conv", 3, filters, filters)
    bn("policy_bn", filters)
    dense("value_fc1", SQUARES, VALUE_HIDDEN)
    conv("policy_out", 1, filters, POLICY_CHANNELS)
    dense("value_fc2", VALUE_HIDDEN, 1)
    return spec


def keras_layers(blocks: int = 7, filters: int = 128, in_planes: int = IN_PLANES):
    """``[(layer_name, [weight_name, ...]), ...]`` covering :func:`weight_spec` in order, with the names tf.keras
    gives the layers of ``network.py:88-120`` (``conv2d``, ``conv2d_1``, ..., ``batch_normalization``, ..., ``dense``,
    ``dense_1``) -- the groups and datasets of a Keras weight file."""
    counters = {"conv2d": 0, "batch_normalization": 0, "dense": 0}
    layers, last = [], None
    for name, shape in weight_spec(blocks, filters, in_planes):
        layer, part = name.rsplit("/", 1)
        if layer != last:
            kind = "batch_normalization" if part == "gamma" else ("conv2d" if len(shape) == 4 else "dense")
            keras_name = kind if counters[kind] == 0 else f"{kind}_{counters[kind]}"
            counters[kind] += 1
            layers.append((keras_name, []))
            last = layer
        layers[-1][1].append(f"{layers[-1][0]}/{part}:0")
    return layers


def num_params(blocks: int = 7, filters: int = 128, in_planes: int = IN_PLANES) -> int:
    return int(sum(int(np.prod(s)) for _, s in weight_spec(blocks, filters, in_planes)))


def _glorot_uniform(rng, shape):
    if len(shape) == 4:
        kh, kw, cin, cout = shape
        fan_in, fan_out = kh * kw * cin, kh * kw * cout
    else:
        fan_in, fan_out = shape
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def init_weights(blocks: int = 7, filters: int = 128, in_planes: int = IN_PLANES,
                 seed: int = 55, perturbed: bool = False):
    """Synthetic weights in ``get_weights()`` layout.

    ``perturbed=False``: Keras defaults.  ``perturbed=True``: random BN statistics and biases so the
    BN fold, bias and residual paths are all exercised by parity tests.
    """
    rng = np.random.default_rng(seed)
    out = []
    for name, shape in weight_spec(blocks, filters, in_planes):
        kind = name.rsplit("/", 1)[1]
        if kind == "kernel":
            w = _glorot_uniform(rng, shape)
        elif kind == "bias":
            w = (rng.normal(0.0, 0.05, size=shape) if perturbed else np.zeros(shape)).astype(np.float32)
        elif kind == "gamma":
            w = (rng.uniform(0.5, 1.5, size=shape) if perturbed else np.ones(shape)).astype(np.float32)
        elif kind == "beta":
            w = (rng.normal(0.0, 0.1, size=shape) if perturbed else np.zeros(shape)).astype(np.float32)
            if perturbed and shape == (1,):
                w[:] = 0.3  # keep the single value-head channel alive so value parity is not vacuous
        elif kind == "moving_mean":
            w = (rng.normal(0.0, 0.1, size=shape) if perturbed else np.zeros(shape)).astype(np.float32)
        elif kind == "moving_variance":
            w = (rng.uniform(0.5, 1.5, size=shape) if perturbed else np.ones(shape)).astype(np.float32)
        else:  # pragma: no cover
            raise AssertionError(name)
        out.append(w)
    return out


def infer_arch(weights):
    """Recover ``(blocks, filters, in_planes)`` from a ``get_weights()`` list."""
    k0 = np.shape(weights[0])
    if len(k0) != 4 or k0[0] != 3 or k0[1] != 3:
        raise ValueError("first tensor must be the 3x3 stem kernel [3,3,in,filters]")
    in_planes, filters = int(k0[2]), int(k0[3])
    n = len(weights)
    # 6 stem + 12*blocks trunk + 18 head tensors
    if (n - 24) % 12 != 0 or n < 24:
        raise ValueError(f"unexpected number of weight tensors: {n}")
    blocks = (n - 24) // 12
    return blocks, filters, in_planes


def adapt_weights(weights, plane_perm=None, plane_scale=None, policy_perm=None):
    """Weights trained on ANOTHER plane / policy-channel layout, re-expressed for this package's layout -- at no run-time
    cost, by permuting the network's first and last tensors.

    The 266 input planes and the 69 policy channels are defined inside minishogilib (``Position.to_alphazero_input``,
    ``MCTS.evaluate``; network.py:258, mcts.py:60,105), which is not part of the reference tree; this package's native
    engine uses its own, documented layout (csrc/minishogi.hpp).  A checkpoint trained by the reference therefore
    matches the native encoder only through a layout table, which a maintainer who has minishogilib can write down:

    * ``plane_perm[c]`` = index, in the checkpoint's layout, of the plane this package calls ``c``; ``plane_scale[c]`` =
      factor between the two conventions (checkpoint's value = scale x ours, e.g. hand counts / 2 against raw counts):
      the stem kernel becomes ``K'[:, :, c, :] = K[:, :, plane_perm[c], :] * plane_scale[c]``;
    * ``policy_perm[k]`` = checkpoint's policy channel for this package's channel ``k`` (logit index ``k*25 + square``):
      the policy 1x1 kernel and bias become ``K'[..., k] = K[..., policy_perm[k]]``.

    Returns a new list; the input is not modified.  ``Network.set_layout_adapter`` applies it on every ``set_weights``."""
    ws = [np.asarray(w, dtype=np.float32) for w in weights]
    _, filters, in_planes = infer_arch(ws)
    out = list(ws)
    if plane_perm is not None or plane_scale is not None:
        perm = np.arange(in_planes) if plane_perm is None else np.asarray(plane_perm, dtype=np.int64)
        scale = np.ones(in_planes, dtype=np.float32) if plane_scale is None else np.asarray(plane_scale, dtype=np.float32)
        if perm.shape != (in_planes,) or scale.shape != (in_planes,) or perm.min() < 0 or perm.max() >= in_planes:
            raise ValueError(f"plane_perm / plane_scale must have {in_planes} entries in [0, {in_planes})")
        out[0] = np.ascontiguousarray(ws[0][:, :, perm, :] * scale.reshape(1, 1, -1, 1))
    if policy_perm is not None:
        perm = np.asarray(policy_perm, dtype=np.int64)
        if perm.shape != (POLICY_CHANNELS,) or sorted(perm.tolist()) != list(range(POLICY_CHANNELS)):
            raise ValueError(f"policy_perm must be a permutation of range({POLICY_CHANNELS})")
        hits = [i for i, w in enumerate(ws) if w.shape == (1, 1, filters, POLICY_CHANNELS)]
        if len(hits) != 1 or ws[hits[0] + 1].shape != (POLICY_CHANNELS,):
            raise ValueError("policy 1x1 convolution not found in the weight list")
        i = hits[0]
        out[i] = np.ascontiguousarray(ws[i][..., perm])
        out[i + 1] = np.ascontiguousarray(ws[i + 1][perm])
    return out
