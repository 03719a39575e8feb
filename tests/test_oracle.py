"""Pins the CPU oracle (test infrastructure) against its own independent restatement, the committed
golden vectors and hand-checkable layout properties.  Reference graph: erweitern_55/network.py:75-146."""
import numpy as np
import pytest
import torch

from erweitern_55_b200 import synth, weights as W
from oracle import network_oracle as O


@pytest.mark.parametrize("tag", ["default", "perturbed"])
def test_torch_numpy_fp64_agree(tag, weights_default, weights_perturbed):
    w = weights_default if tag == "default" else weights_perturbed
    x = synth.synthetic_planes(3, seed=7)
    p64, v64 = O.forward_torch(w, x, torch.float64)
    pn, vn = O.forward_numpy(w, x)
    assert p64.shape == (3, 1725) and v64.shape == (3, 1)
    np.testing.assert_allclose(p64, pn, atol=1e-11)
    np.testing.assert_allclose(v64, vn, atol=1e-12)
    p32, v32 = O.forward_torch(w, x, torch.float32)
    assert p32.dtype == np.float32 and v32.dtype == np.float32
    assert np.abs(p32 - p64).max() < 2e-5 and np.abs(v32 - v64).max() < 2e-6
    assert np.all(np.abs(v64) < 1.0)


@pytest.mark.parametrize("tag", ["default", "perturbed"])
def test_golden_vectors(tag, golden, weights_default, weights_perturbed):
    w = weights_default if tag == "default" else weights_perturbed
    np.testing.assert_allclose([t.astype(np.float64).sum() for t in w], golden[f"{tag}_weight_sums"], rtol=0, atol=1e-9)
    x = synth.synthetic_planes(int(golden["batch"]), seed=int(golden["input_seed"]))
    assert x.astype(np.float64).sum() == float(golden["input_checksum"])
    p64, v64 = O.forward_torch(w, x, torch.float64)
    np.testing.assert_allclose(p64, golden[f"{tag}_policy_f64"], atol=1e-10)
    np.testing.assert_allclose(v64, golden[f"{tag}_value_f64"], atol=1e-10)
    pb, vb = O.forward_bf16_emulated(w, x)
    np.testing.assert_allclose(pb, golden[f"{tag}_policy_bf16"], atol=2e-4)
    np.testing.assert_allclose(vb, golden[f"{tag}_value_bf16"], atol=2e-4)


def test_bn_fold_equivalence(weights_perturbed):
    net = O.parse_weights(weights_perturbed)
    rng = np.random.default_rng(3)
    x = rng.random((2, 128, 5, 5))
    blk = net["blocks"][0]
    ref = O._bn_np(O._conv_np(x, blk["conv1"]), blk["bn1"])
    k, b = O.fold_conv_bn(blk["conv1"], blk["bn1"])
    folded = O._conv_np(x, {"kernel": k, "bias": b})
    np.testing.assert_allclose(folded, ref, atol=1e-12)


def test_bf16_emulation_within_stated_tolerance(weights_default):
    """bf16 path |dlogit| <= 0.03, |dvalue| <= 0.01 against the fp32 graph."""
    x = synth.synthetic_planes(32, seed=11)
    p, v = O.forward_torch(weights_default, x)
    pb, vb = O.forward_bf16_emulated(weights_default, x)
    assert np.abs(p - pb).max() <= 0.03
    assert np.abs(v - vb).max() <= 0.01


def test_same_padding_and_flatten_order():
    """Impulse through an identity-like net: checks 'same' padding, tap orientation and c*25+y*5+x."""
    w = W.init_weights(blocks=0, filters=128, in_planes=4, seed=1)
    spec = W.weight_spec(blocks=0, filters=128, in_planes=4)
    names = [n for n, _ in spec]
    w = [np.zeros_like(t) for t in w]
    for part in ("gamma", "moving_variance"):
        for i, n in enumerate(names):
            if n.endswith(part):
                w[i][:] = 1.0 - (1e-3 if part == "moving_variance" else 0.0)  # scale exactly 1
    # stem: output channel 0 copies input channel 2 shifted: out(y,x) = in(y+1, x-1)  (tap ky=2,kx=0)
    w[names.index("stem_conv/kernel")][2, 0, 2, 0] = 1.0
    # policy conv 3x3: centre tap identity on channel 0
    w[names.index("policy_conv/kernel")][1, 1, 0, 0] = 1.0
    # policy out: channel 5 = 3 * channel 0, bias 0.25 on channel 7
    w[names.index("policy_out/kernel")][0, 0, 0, 5] = 3.0
    w[names.index("policy_out/bias")][7] = 0.25
    x = np.zeros((1, 4, 5, 5), dtype=np.float32)
    x[0, 2, 3, 1] = 1.0     # in(y=3, x=1) -> out(y=2, x=2)
    x[0, 2, 0, 0] = 1.0     # in(0,0) -> out(-1,1): falls off the board
    for fwd in (lambda: O.forward_torch(w, x, torch.float64), lambda: O.forward_numpy(w, x)):
        p, _ = fwd()
        expect = np.zeros(1725)
        expect[5 * 25 + 2 * 5 + 2] = 3.0
        expect[7 * 25:8 * 25] = 0.25
        np.testing.assert_allclose(p[0], expect, atol=1e-12)


def test_masked_softmax_oracle():
    rng = np.random.default_rng(0)
    p = rng.normal(size=(3, 1725))
    m = synth.synthetic_legal_mask(3, seed=5)
    m[2] = 0
    pr = O.masked_softmax(p, m)
    assert np.allclose(pr[:2].sum(1), 1.0) and np.all(pr[2] == 0) and np.all(pr[m == 0] == 0)
    i = np.flatnonzero(m[0])
    np.testing.assert_allclose(pr[0, i], np.exp(p[0, i]) / np.exp(p[0, i]).sum(), rtol=1e-12)


def test_tf_golden_pins_the_oracle():
    """When tests/golden/tf_golden.npz exists (written by tools/make_tf_golden.py where the reference's TensorFlow 1.15
    stack runs), the oracle must reproduce the reference's own predict() outputs: that is what pins it."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "tf_golden.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/tf_golden.npz not present: the forward-pass oracle stays 'parity unpinned'")
    z = np.load(path)
    w = [z[f"w{i}"] for i in range(int(z["n_weights"]))]
    from oracle import network_oracle as O
    p, v = O.forward_torch(w, z["inputs"])
    assert np.abs(p - z["policy"]).max() <= 1e-4 and np.abs(v - z["value"]).max() <= 1e-5


def test_make_tf_golden_generators_match_this_repo():
    """tools/make_tf_golden.py carries its own copy of the synthetic-input generator (it must run without this package):
    the copy and the original produce the same planes, and its weight perturbation keeps the reference's shapes."""
    import importlib.util
    import os
    from erweitern_55_b200 import synth, weights as W
    spec = importlib.util.spec_from_file_location("make_tf_golden", os.path.join(os.path.dirname(__file__), "..", "tools", "make_tf_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert np.array_equal(mod.synthetic_planes(7, seed=1007), synth.synthetic_planes(7, seed=1007))
    w = W.init_weights(seed=1)
    pw = mod.perturb(w, 55)
    assert [a.shape for a in pw] == [a.shape for a in w] and W.infer_arch(pw) == (7, 128, 266)
    from oracle import network_oracle as O
    p, v = O.forward_torch(pw, synth.synthetic_planes(3, seed=2))
    assert np.isfinite(p).all() and np.abs(v).max() < 1


def test_layout_adapter_algebra():
    """weights.adapt_weights: a network trained on another plane order / plane scaling / policy-channel order, re-expressed
    for this package's layout, gives the same outputs on correspondingly re-laid-out inputs (oracle arithmetic, fp64)."""
    rng = np.random.default_rng(11)
    w = W.init_weights(seed=3, perturbed=True)
    plane_perm = rng.permutation(266)
    plane_scale = rng.choice([1.0, 0.5, 2.0], size=266).astype(np.float32)
    policy_perm = rng.permutation(69)
    x_ours = synth.synthetic_planes(6, seed=9)
    x_ref = np.zeros_like(x_ours)
    x_ref[:, plane_perm] = x_ours * plane_scale.reshape(1, -1, 1, 1)          # what the checkpoint's own encoder would have produced
    p_ref, v_ref = O.forward_torch(w, x_ref, dtype=torch.float64)
    p_ours, v_ours = O.forward_torch(W.adapt_weights(w, plane_perm, plane_scale, policy_perm), x_ours, dtype=torch.float64)
    assert np.abs(v_ours - v_ref).max() < 1e-6                                # (the adapted list is float32: 1e-7 rounding of scaled kernels)
    want = p_ref.reshape(6, 69, 25)[:, policy_perm].reshape(6, 1725)          # our channel k = the checkpoint's channel policy_perm[k]
    assert np.abs(p_ours - want).max() < 1e-5
    assert all(a is b or np.array_equal(a, b) for a, b in zip(w, W.adapt_weights(w)))
    with pytest.raises(ValueError):
        W.adapt_weights(w, policy_perm=np.zeros(69, dtype=int))
