"""CPU oracle for the erweitern_55 policy/value forward pass.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``erweitern_55_b200``) never imports anything under ``oracle/``.

PARITY UNPINNED.  The reference executes this graph inside TensorFlow 1.15.2 / Keras
(``requirements.txt:32``), which is not vendored in ``/root/reference`` and cannot be installed here
(no wheel, no network, no Python-3.6/3.7).  The reference ships no test, golden vector or fixture for
the forward pass (``tests/test_checkmate.py`` and ``tests/test_repetition.py`` only touch minishogilib).
The oracle is therefore a restatement of the published Keras layer semantics, anchored on the
reference's own graph definition:

  * graph:           ``erweitern_55/network.py:75-122``  (``_alphazero_network``)
  * residual block:  ``erweitern_55/network.py:124-146`` (ReLU after the add, none before it)
  * predict:         ``erweitern_55/network.py:201-216`` (whole input is one batch; returns policy, value)
  * weight list:     ``erweitern_55/network.py:228-243`` (Keras ``get_weights`` order)

Keras-1.15 defaults hard-coded here: Conv2D ``use_bias=True`` with HWIO kernels, ``padding='same'``
(= 1 zero on each side for 3x3 stride 1), ``data_format='channels_first'``; BatchNormalization(axis=1)
in inference mode ``y = gamma * (x - moving_mean) / sqrt(moving_variance + 1e-3) + beta``; Dense kernels
``[in, out]``; ``Flatten`` of an NCHW tensor keeps memory order, so policy index = c*25 + y*5 + x
(matches the label reshape ``[B, 69*5*5]`` at ``train.py:71``).

Two independent restatements live here and are cross-checked by ``tests/test_oracle.py``:
``forward_torch`` (torch CPU conv2d, fp32 or fp64) and ``forward_numpy`` (explicit zero padding +
einsum, fp64).  ``forward_bf16_emulated`` restates the numerics of the tensor-core kernel
(BN folded in fp64, weights and inter-layer activations rounded to bf16, fp32 accumulation) so that the
GPU path can be checked tightly as well as against the fp32 graph.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
SQUARES = 25
POLICY_CHANNELS = 69


# --------------------------------------------------------------------------------------------------
# weight list parsing (canonical Keras order; network.py:228-243)
# --------------------------------------------------------------------------------------------------
def parse_weights(weights):
    """Split a ``get_weights()`` list into named layers.  Asserts the canonical order and shapes."""
    ws = [np.asarray(w) for w in weights]
    n = len(ws)
    assert (n - 24) % 12 == 0 and n >= 24, f"unexpected tensor count {n}"
    blocks = (n - 24) // 12
    it = iter(ws)

    def conv(kh, cout=None):
        k, b = next(it), next(it)
        assert k.ndim == 4 and k.shape[0] == kh and k.shape[1] == kh, k.shape
        assert b.shape == (k.shape[3],)
        if cout is not None:
            assert k.shape[3] == cout, (k.shape, cout)
        return {"kernel": k, "bias": b}

    def bn(c):
        g, b, m, v = next(it), next(it), next(it), next(it)
        for t in (g, b, m, v):
            assert t.shape == (c,), (t.shape, c)
        return {"gamma": g, "beta": b, "mean": m, "var": v}

    def dense(cin, cout):
        k, b = next(it), next(it)
        assert k.shape == (cin, cout) and b.shape == (cout,), (k.shape, b.shape)
        return {"kernel": k, "bias": b}

    net = {"blocks": []}
    net["stem_conv"] = conv(3)
    filters = net["stem_conv"]["kernel"].shape[3]
    net["stem_bn"] = bn(filters)
    for _ in range(blocks):
        blk = {"conv1": conv(3, filters)}
        blk["bn1"] = bn(filters)
        blk["conv2"] = conv(3, filters)
        blk["bn2"] = bn(filters)
        net["blocks"].append(blk)
    net["value_conv"] = conv(1, 1)
    net["value_bn"] = bn(1)
    net["policy_conv"] = conv(3, filters)
    net["policy_bn"] = bn(filters)
    net["value_fc1"] = dense(SQUARES, 128)
    net["policy_out"] = conv(1, POLICY_CHANNELS)
    net["value_fc2"] = dense(128, 1)
    assert next(it, None) is None
    return net


# --------------------------------------------------------------------------------------------------
# restatement 1: torch CPU (fp32 = the reference's arithmetic type; fp64 twin bounds its rounding)
# --------------------------------------------------------------------------------------------------
def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _conv_t(x, layer, dtype):
    k = _t(layer["kernel"], dtype).permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
    pad = k.shape[2] // 2                                             # 'same'
    return F.conv2d(x, k, _t(layer["bias"], dtype), stride=1, padding=pad)


def _bn_t(x, layer, dtype):
    g, b = _t(layer["gamma"], dtype), _t(layer["beta"], dtype)
    m, v = _t(layer["mean"], dtype), _t(layer["var"], dtype)
    scale = g / torch.sqrt(v + BN_EPS)
    return (x - m.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def prepare_torch(weights, dtype=torch.float32):
    """Pre-convert a weight list once (used by the timed CPU baseline)."""
    net = parse_weights(weights)

    def cv(layer):
        k = _t(layer["kernel"], dtype).permute(3, 2, 0, 1).contiguous()
        return k, _t(layer["bias"], dtype), k.shape[2] // 2

    def bnp(layer):
        g, b = _t(layer["gamma"], dtype), _t(layer["beta"], dtype)
        m, v = _t(layer["mean"], dtype), _t(layer["var"], dtype)
        scale = (g / torch.sqrt(v + BN_EPS)).view(1, -1, 1, 1)
        shift = b.view(1, -1, 1, 1) - m.view(1, -1, 1, 1) * scale
        return scale, shift

    prep = {"dtype": dtype, "stem": (cv(net["stem_conv"]), bnp(net["stem_bn"])), "blocks": []}
    for blk in net["blocks"]:
        prep["blocks"].append((cv(blk["conv1"]), bnp(blk["bn1"]), cv(blk["conv2"]), bnp(blk["bn2"])))
    prep["policy"] = (cv(net["policy_conv"]), bnp(net["policy_bn"]), cv(net["policy_out"]))
    prep["value"] = (cv(net["value_conv"]), bnp(net["value_bn"]),
                     _t(net["value_fc1"]["kernel"], dtype), _t(net["value_fc1"]["bias"], dtype),
                     _t(net["value_fc2"]["kernel"], dtype), _t(net["value_fc2"]["bias"], dtype))
    return prep


def forward_prepared(prep, x, return_trunk=False):
    """Forward pass on a prepared weight set; ``x`` torch tensor ``[N,C,5,5]``.  network.py:87-122.
    return_trunk: also the trunk output ``[N,filters,5,5]`` (the input of both heads, network.py:99)."""
    def cb(x, c, b):
        y = F.conv2d(x, c[0], c[1], stride=1, padding=c[2])
        return y * b[0] + b[1]

    with torch.no_grad():
        x = x.to(prep["dtype"])
        h = F.relu(cb(x, *prep["stem"]))                                  # network.py:92-95
        for c1, b1, c2, b2 in prep["blocks"]:                             # network.py:98-99,124-146
            y = F.relu(cb(h, c1, b1))
            y = cb(y, c2, b2)
            h = F.relu(y + h)
        pc, pb, po = prep["policy"]                                       # network.py:102-108
        p = F.relu(cb(h, pc, pb))
        p = F.conv2d(p, po[0], po[1])
        policy = p.reshape(p.shape[0], -1)
        vc, vb, k1, b1_, k2, b2_ = prep["value"]                          # network.py:111-120
        v = F.relu(cb(h, vc, vb)).reshape(h.shape[0], -1)
        v = F.relu(v @ k1 + b1_)
        value = torch.tanh(v @ k2 + b2_)
    if return_trunk:
        return policy, value, h
    return policy, value


def forward_torch(weights, images, dtype=torch.float32, return_trunk=False):
    """``Network.predict`` restated: ``images [N,266,5,5]`` -> ``(policy [N,1725], value [N,1])`` numpy
    (and the trunk output with ``return_trunk``)."""
    prep = prepare_torch(weights, dtype)
    x = torch.from_numpy(np.ascontiguousarray(np.asarray(images))).to(dtype)
    out = forward_prepared(prep, x, return_trunk)
    keep = torch.float64 if dtype == torch.float64 else torch.float32
    return tuple(t.to(keep).numpy() for t in out)


# --------------------------------------------------------------------------------------------------
# restatement 2: plain numpy fp64 (explicit padding, one einsum per tap) -- independent of torch conv
# --------------------------------------------------------------------------------------------------
def _conv_np(x, layer):
    k = layer["kernel"].astype(np.float64)            # [kh,kw,ci,co]
    kh = k.shape[0]
    pad = kh // 2
    n, _, h, w = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    out = np.zeros((n, k.shape[3], h, w), dtype=np.float64)
    for dy in range(kh):
        for dx in range(kh):
            out += np.einsum("ncyx,co->noyx", xp[:, :, dy:dy + h, dx:dx + w], k[dy, dx])
    return out + layer["bias"].astype(np.float64).reshape(1, -1, 1, 1)


def _bn_np(x, layer):
    g, b = layer["gamma"].astype(np.float64), layer["beta"].astype(np.float64)
    m, v = layer["mean"].astype(np.float64), layer["var"].astype(np.float64)
    r = lambda t: t.reshape(1, -1, 1, 1)
    return r(g) * (x - r(m)) / np.sqrt(r(v) + BN_EPS) + r(b)


def forward_numpy(weights, images):
    net = parse_weights(weights)
    relu = lambda t: np.maximum(t, 0.0)
    x = np.asarray(images, dtype=np.float64)
    h = relu(_bn_np(_conv_np(x, net["stem_conv"]), net["stem_bn"]))
    for blk in net["blocks"]:
        y = relu(_bn_np(_conv_np(h, blk["conv1"]), blk["bn1"]))
        y = _bn_np(_conv_np(y, blk["conv2"]), blk["bn2"])
        h = relu(y + h)
    p = relu(_bn_np(_conv_np(h, net["policy_conv"]), net["policy_bn"]))
    p = _conv_np(p, net["policy_out"])
    policy = p.reshape(p.shape[0], -1)
    v = relu(_bn_np(_conv_np(h, net["value_conv"]), net["value_bn"])).reshape(h.shape[0], -1)
    v = relu(v @ net["value_fc1"]["kernel"].astype(np.float64) + net["value_fc1"]["bias"])
    value = np.tanh(v @ net["value_fc2"]["kernel"].astype(np.float64) + net["value_fc2"]["bias"])
    return policy, value


# --------------------------------------------------------------------------------------------------
# BN folding + bf16 emulation of the tensor-core path
# --------------------------------------------------------------------------------------------------
def fold_conv_bn(conv, bn=None):
    """Fold inference BN into the preceding conv (fp64 math): returns (kernel HWIO f64, bias f64)."""
    k = conv["kernel"].astype(np.float64)
    b = conv["bias"].astype(np.float64)
    if bn is None:
        return k, b
    scale = bn["gamma"].astype(np.float64) / np.sqrt(bn["var"].astype(np.float64) + BN_EPS)
    return k * scale.reshape(1, 1, 1, -1), (b - bn["mean"].astype(np.float64)) * scale + bn["beta"].astype(np.float64)


def bf16_round(a):
    """Round-to-nearest-even to bfloat16, returned as float32."""
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
    return t.to(torch.bfloat16).to(torch.float32).numpy()


def forward_bf16_emulated(weights, images, return_trunk=False):
    """Numerics of the tcgen05 path: folded bf16 weights, bf16 inter-layer activations, fp32 accumulate.

    Every 3x3 conv and the policy 1x1 conv take bf16 operands (products exact in fp32) with fp32
    accumulation; biases are fp32; the residual operand is the stored bf16 block input; the value head
    runs in fp32 on the bf16 trunk output.  Accumulation order differs from the kernel's, so expect
    agreement to ~1e-3 on logits rather than bit-exactness.
    """
    net = parse_weights(weights)

    def conv(x, layer, bn):
        k, b = fold_conv_bn(layer, bn)
        k = torch.from_numpy(bf16_round(k.astype(np.float32))).permute(3, 2, 0, 1).contiguous()
        return F.conv2d(x, k, torch.from_numpy(b.astype(np.float32)), padding=k.shape[2] // 2)

    rb = lambda t: t.to(torch.bfloat16).to(torch.float32)
    with torch.no_grad():
        x = rb(torch.from_numpy(np.ascontiguousarray(np.asarray(images, dtype=np.float32))))
        h = rb(F.relu(conv(x, net["stem_conv"], net["stem_bn"])))
        for blk in net["blocks"]:
            y = rb(F.relu(conv(h, blk["conv1"], blk["bn1"])))
            h = rb(F.relu(conv(y, blk["conv2"], blk["bn2"]) + h))
        p = rb(F.relu(conv(h, net["policy_conv"], net["policy_bn"])))
        policy = conv(p, net["policy_out"], None).reshape(p.shape[0], -1)
        kv, bv = fold_conv_bn(net["value_conv"], net["value_bn"])
        kv = torch.from_numpy(kv.astype(np.float32)).permute(3, 2, 0, 1).contiguous()
        v = F.relu(F.conv2d(h, kv, torch.from_numpy(bv.astype(np.float32)))).reshape(h.shape[0], -1)
        v = F.relu(v @ torch.from_numpy(net["value_fc1"]["kernel"].astype(np.float32))
                   + torch.from_numpy(net["value_fc1"]["bias"].astype(np.float32)))
        value = torch.tanh(v @ torch.from_numpy(net["value_fc2"]["kernel"].astype(np.float32))
                           + torch.from_numpy(net["value_fc2"]["bias"].astype(np.float32)))
    if return_trunk:
        return policy.numpy(), value.numpy(), h.numpy()
    return policy.numpy(), value.numpy()


def masked_softmax(policy, legal_mask):
    """Oracle for the fused policy tail: ``softmax(where(mask, logits, -inf))`` per row (fp64).

    The reference leaves masking + softmax to ``minishogilib.MCTS.evaluate`` (call site ``mcts.py:60,105-106``).
    Rows without any legal move return all zeros.
    """
    z = np.where(np.asarray(legal_mask).astype(bool), np.asarray(policy, dtype=np.float64), -np.inf)
    m = np.max(z, axis=1, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    e = np.exp(z - m)
    s = e.sum(axis=1, keepdims=True)
    return np.where(s > 0, e / np.where(s > 0, s, 1.0), 0.0)
