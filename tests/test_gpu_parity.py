"""GPU parity tests proper: the CUDA paths, called through the C ABI (ctypes), against the CPU oracle.

Tolerances (BASELINE.md section 4 / BASELINE.json north_star):
  fp32 CUDA-core path  : |dlogit| <= 1e-4, |dvalue| <= 1e-5 against the fp32 oracle
  bf16 tcgen05 path    : |dlogit| <= 0.03, |dvalue| <= 0.01 against the fp32 oracle, identical top-1
                         wherever the oracle's top-2 margin exceeds 2 x the logit tolerance.
                         bf16 rounding error scales with the logit magnitude: the "perturbed" weight set
                         (random BN statistics) doubles the logit range (+-4 instead of +-2), so its logit
                         tolerance is 0.06.
"""
import threading

import numpy as np
import pytest
import torch

from erweitern_55_b200 import synth, weights as W

pytestmark = pytest.mark.gpu

FP32_TOL = (1e-4, 1e-5)
BF16_TOL = (0.03, 0.01)
BF16_TOL_PERTURBED = (0.06, 0.01)


def _bf16_tol(tag):
    return BF16_TOL if tag == "default" else BF16_TOL_PERTURBED


def _oracle():
    from oracle import network_oracle as O
    return O


@pytest.fixture(scope="module")
def nets(weights_default, weights_perturbed):
    from erweitern_55_b200.network import Network
    made = {}
    for tag, w in (("default", weights_default), ("perturbed", weights_perturbed)):
        for prec in ("fp32", "bf16"):
            made[tag, prec] = Network(device="gpu", precision=prec, weights=w)
    yield made
    for n in made.values():
        n.close()


def _check(p, v, po, vo, tol, min_decisive=None, label=""):
    """Outputs against the oracle's: absolute tolerances, and the same top-1 move wherever the oracle's top-2 margin
    exceeds twice the logit tolerance ("decisive" rows).  The share of decisive rows is asserted (min_decisive; from 16
    rows on, by default at least 0.3: random-init networks have a median top-2 margin of 0.088 against the 0.06
    threshold of the bf16 tolerance, BASELINE.md section 4, so 40-60 % of their rows are decisive; the trained-like set
    asks for 0.5 and has ~0.9) and the tie rate -- rows too close to call -- is reported (pytest -s shows it).  From 64
    rows on, top-1 must also agree on at least 95 % of ALL rows, near-ties included (BASELINE.md section 4 measured
    99.1 % over 2 048 rows for bf16 rounding as such; 64 rows can lose two)."""
    assert p.shape == po.shape and v.shape == vo.shape and p.dtype == np.float32 and v.dtype == np.float32
    assert np.isfinite(p).all() and np.isfinite(v).all()
    assert np.abs(p - po).max() <= tol[0], np.abs(p - po).max()
    assert np.abs(v - vo).max() <= tol[1], np.abs(v - vo).max()
    top2 = np.sort(po, axis=1)[:, -2:]
    decisive = (top2[:, 1] - top2[:, 0]) > 2 * tol[0]
    assert (p.argmax(1)[decisive] == po.argmax(1)[decisive]).all()
    frac = float(decisive.mean())
    if min_decisive is None:
        min_decisive = 0.3 if len(p) >= 16 else 0.0
    if len(p) >= 64:
        assert float((p.argmax(1) == po.argmax(1)).mean()) >= 0.95
    print(f"[parity{' ' + label if label else ''}] {len(p)} rows: max |dlogit| {np.abs(p - po).max():.2e}, max |dvalue| {np.abs(v - vo).max():.2e}, "
          f"decisive {frac:.3f} (tie rate {1 - frac:.3f}), top-1 equal on all rows {float((p.argmax(1) == po.argmax(1)).mean()):.3f}")
    assert frac >= min_decisive, (frac, min_decisive)
    return frac


@pytest.mark.parametrize("tag", ["default", "perturbed"])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 16, 37, 64, 255])
def test_predict_matches_oracle(nets, weights_default, weights_perturbed, tag, prec, n):
    w = weights_default if tag == "default" else weights_perturbed
    x = synth.synthetic_planes(n, seed=1000 + n)
    p, v = nets[tag, prec].predict(x)
    po, vo = _oracle().forward_torch(w, x)
    _check(p, v, po, vo, FP32_TOL if prec == "fp32" else _bf16_tol(tag), label=f"{tag} {prec} batch {n}")


@pytest.mark.parametrize("tag", ["default", "perturbed"])
def test_golden_fixture(nets, golden, tag):
    x = synth.synthetic_planes(int(golden["batch"]), seed=int(golden["input_seed"]))
    p, v = nets[tag, "fp32"].predict(x)
    _check(p, v, golden[f"{tag}_policy_f64"].astype(np.float32), golden[f"{tag}_value_f64"].astype(np.float32), FP32_TOL, label=f"golden {tag} fp32")
    p, v = nets[tag, "bf16"].predict(x)
    _check(p, v, golden[f"{tag}_policy_f64"].astype(np.float32), golden[f"{tag}_value_f64"].astype(np.float32), _bf16_tol(tag), label=f"golden {tag} bf16")
    # the oracle's bf16 emulation (same rounding points, different accumulation order) stays as close
    assert np.abs(p - golden[f"{tag}_policy_bf16"]).max() <= _bf16_tol(tag)[0]


def test_trunk_activations_match_the_oracle(nets, weights_perturbed):
    """Per-layer parity against the CPU oracle, not against another CUDA path: the trunk output (after all 15 3x3
    convolutions) of both precision paths against the oracle's fp32 trunk, and the tcgen05 path against the oracle's
    bf16 emulation (same rounding points: folded bf16 weights, bf16 activations between layers, fp32 accumulation);
    intermediate layers of the two CUDA paths are held against each other as before."""
    O = _oracle()
    x = synth.synthetic_planes(11, seed=3)
    _, _, trunk_emulated = O.forward_bf16_emulated(weights_perturbed, x, return_trunk=True)
    _, _, trunk_fp32 = O.forward_torch(weights_perturbed, x, return_trunk=True)
    a32 = nets["perturbed", "fp32"].debug_activations(x)
    a16 = nets["perturbed", "bf16"].debug_activations(x)
    scale = max(1.0, float(np.abs(trunk_fp32).max()))
    assert np.abs(a32 - trunk_fp32).max() <= 1e-4 * scale
    assert np.abs(a16 - trunk_fp32).max() <= 0.02 * scale
    # against the emulation only accumulation order differs (a sum that lands on the other side of a bf16 rounding
    # boundary moves an activation by one bf16 step, 2^-8 relative, and the layers after it follow)
    assert np.abs(a16 - trunk_emulated).max() <= 0.008 * scale, np.abs(a16 - trunk_emulated).max()
    assert np.mean(np.abs(a16 - trunk_emulated)) <= 5e-4 * scale, np.mean(np.abs(a16 - trunk_emulated))
    for nconv in (1, 2, 3):
        a = nets["perturbed", "bf16"].debug_activations(x, nconv)
        b = nets["perturbed", "fp32"].debug_activations(x, nconv)
        assert np.abs(a - b).max() <= 0.02 * max(1.0, np.abs(b).max())


def test_reference_call_conventions(nets):
    """What mcts.py does with the result (mcts.py:57-60,101-106) and the float64 warm-up (network.py:72)."""
    net = nets["default", "bf16"]
    x64 = np.random.default_rng(0).random((1, 266, 5, 5))           # float64, batch 1
    policy, value = net.predict(x64)
    assert policy.shape == (1, 1725) and value.shape == (1, 1)
    v = (value + 1) / 2
    assert 0.0 <= float(v[0][0]) <= 1.0 and policy[0].shape == (1725,)
    plist, vlist = net.model.predict(x64.tolist())                     # array-like input
    np.testing.assert_array_equal(plist, policy)
    with pytest.raises(ValueError):
        net.predict(np.zeros((0, 266, 5, 5), dtype=np.float32))
    with pytest.raises(ValueError):
        net.predict(np.zeros((2, 265, 5, 5), dtype=np.float32))
    z, vz = net.predict(net.zero_inputs(3))                            # all-zero planes: bias-only path
    assert np.allclose(z[0], z[1]) and np.allclose(vz[0], vz[2])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_batch_invariance_and_determinism(nets, prec):
    """Size-independent properties at the headline batch: row i of a 1024 batch == the same position
    evaluated alone or in another order; repeated calls are bit-identical."""
    net = nets["perturbed", prec]
    x = synth.synthetic_planes(1024, seed=77)
    p, v = net.predict(x)
    p2, v2 = net.predict(x)
    assert np.array_equal(p, p2) and np.array_equal(v, v2)
    perm = np.random.default_rng(1).permutation(1024)
    pp, vp = net.predict(x[perm])
    tol = 1e-5 if prec == "fp32" else 0.0   # tcgen05 path: a position's arithmetic does not depend on its slot
    assert np.abs(pp - p[perm]).max() <= tol and np.abs(vp - v[perm]).max() <= tol
    for i in (0, 5, 1023):
        pi, vi = net.predict(x[i:i + 1])
        assert np.abs(pi[0] - p[i]).max() <= max(tol, 1e-6) and abs(vi[0, 0] - v[i, 0]) <= max(tol, 1e-6)


def test_large_batch_against_oracle_sample(nets, weights_default):
    net = nets["default", "bf16"]
    x = synth.synthetic_planes(4096, seed=9)
    p, v = net.predict(x)
    idx = np.arange(0, 4096, 97)
    po, vo = _oracle().forward_torch(weights_default, x[idx])
    _check(p[idx], v[idx], po, vo, BF16_TOL, label="4096 sample")


@pytest.mark.parametrize("tag", ["default", "perturbed"])
@pytest.mark.parametrize("n", [1024, 1184])
def test_headline_batches_against_the_oracle(nets, weights_default, weights_perturbed, tag, n):
    """Every row of the headline batch (1024) and of the full-machine batch (1184 = 148 SMs x 8 positions) against the
    oracle directly, through Network.predict (host pipeline: packed chunks) and through the device-resident call."""
    w = weights_default if tag == "default" else weights_perturbed
    x = synth.synthetic_planes(n, seed=4000 + n)
    po, vo = _oracle().forward_torch(w, x)
    net = nets[tag, "bf16"]
    p, v = net.predict(x)
    _check(p, v, po, vo, _bf16_tol(tag), label=f"{tag} batch {n}")
    xd = torch.from_numpy(x).cuda()
    pd, vd = torch.empty((n, 1725), device="cuda"), torch.empty((n, 1), device="cuda")
    net.predict_device(xd, pd, vd)
    net.sync()
    assert np.array_equal(pd.cpu().numpy(), p) and np.array_equal(vd.cpu().numpy(), v)   # one launch == the chunked pipeline


def test_trained_like_dynamic_range(weights_trained_like):
    """bf16 tolerance on a network with a trained network's dynamic range (tests/conftest.py::calibrate_trained_like:
    BatchNorm statistics far from the identity, logits from -14 to +20 instead of +-2): the rounding error of bf16
    operands scales with the magnitudes, so the tolerance is relative -- 1.5 % of the largest |logit| of the batch
    (measured: 1.0 % for the oracle's own bf16 emulation) -- and top-1 must agree wherever the oracle's margin exceeds
    twice that.  The fp32 path stays at 1e-4 relative."""
    from erweitern_55_b200.network import Network
    O = _oracle()
    x = synth.synthetic_planes(512, seed=5)
    po, vo = O.forward_torch(weights_trained_like, x)
    scale = float(np.abs(po).max())
    assert scale > 10.0
    net = Network(precision="fp32", weights=weights_trained_like)
    p, v = net.predict(x[:64])
    _check(p, v, po[:64], vo[:64], (1e-4 * scale, 1e-4), label="trained-like fp32")
    net.set_precision("bf16")
    p, v = net.predict(x)
    frac = _check(p, v, po, vo, (0.015 * scale, 0.05), min_decisive=0.5, label="trained-like bf16")
    pe, ve = O.forward_bf16_emulated(weights_trained_like, x)
    assert np.abs(p - pe).max() <= 0.008 * scale and np.abs(v - ve).max() <= 0.02      # same rounding points, other summation order
    assert (p.argmax(1) == po.argmax(1)).mean() >= 0.95
    net.close()


def test_tf_golden_when_present(nets):
    """Reference-side pin: tests/golden/tf_golden.npz is written by tools/make_tf_golden.py on a machine that has the
    reference's TensorFlow 1.15 / Keras stack (it cannot be produced here) -- the reference's own Network with seeded
    weights, model.get_weights() and predict() on the seeded synthetic inputs.  When the file is there, both CUDA
    paths are held to the reference's outputs themselves, and the oracle with them."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "tf_golden.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/tf_golden.npz not present (produce it with tools/make_tf_golden.py on a TF-1.15 machine)")
    from erweitern_55_b200.network import Network
    z = np.load(path)
    w = [z[f"w{i}"] for i in range(int(z["n_weights"]))]
    x, po, vo = z["inputs"], z["policy"], z["value"]
    p_or, v_or = _oracle().forward_torch(w, x)
    assert np.abs(p_or - po).max() <= 1e-4 and np.abs(v_or - vo).max() <= 1e-5      # the oracle IS the reference's forward pass
    net = Network(precision="fp32", weights=w)
    _check(*net.predict(x), po, vo, FP32_TOL, label="TF golden fp32")
    net.set_precision("bf16")
    _check(*net.predict(x), po, vo, BF16_TOL, label="TF golden bf16")
    net.close()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_masked_softmax_tail(nets, weights_default, prec):
    O = _oracle()
    n = 19
    x = synth.synthetic_planes(n, seed=21)
    m = synth.synthetic_legal_mask(n, seed=22)
    m[3] = 0                                  # a row without legal moves
    m[4] = 0; m[4, 100] = 1                   # a single legal move
    probs, v = nets["default", prec].predict_masked(x, m)
    p, v2 = nets["default", prec].predict(x)
    np.testing.assert_array_equal(v, v2)
    ref = O.masked_softmax(p, m)              # softmax of the SAME logits: tests the fused tail alone
    assert np.abs(probs - ref).max() <= 2e-6
    assert np.all(probs[m == 0] == 0) and np.all(probs[3] == 0) and probs[4, 100] == 1.0
    po, _ = O.forward_torch(weights_default, x)
    assert np.abs(probs - O.masked_softmax(po, m)).max() <= (1e-4 if prec == "fp32" else 0.03)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("n", [1, 9, 300, 1024])
def test_packed_input_is_identical_to_float_planes(nets, prec, n):
    """e55_predict_packed must reproduce e55_predict bit for bit (same kernel arithmetic, different input format)."""
    from erweitern_55_b200 import packing
    net = nets["perturbed", prec]
    x = synth.synthetic_planes(n, seed=50 + n)
    bits, scale = packing.pack_planes(x)
    p, v = net.predict(x)
    pp, vp = net.predict_packed(bits, scale)
    assert np.array_equal(p, pp) and np.array_equal(v, vp)
    with pytest.raises(ValueError):
        net.predict_packed(bits[:, :10], scale[:, :10])


def test_pickled_weight_wire_format(nets, weights_perturbed):
    """The client/trainer exchange weights as a pickled get_weights() list (train.py:50, client.py:47-51)."""
    import pickle
    from erweitern_55_b200.network import Network
    blob = pickle.dumps(weights_perturbed, protocol=4)
    net = Network(precision="bf16", seed=1)
    net.load_pickled_weights(blob)
    x = synth.synthetic_planes(5, seed=77)
    assert np.array_equal(net.predict(x)[0], nets["perturbed", "bf16"].predict(x)[0])
    again = pickle.loads(net.dump_pickled_weights())
    assert all(np.array_equal(a, b) for a, b in zip(again, weights_perturbed))
    net.close()


def test_set_weights_roundtrip_and_swap(nets, weights_default, weights_perturbed, tmp_path):
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_default)
    x = synth.synthetic_planes(6, seed=4)
    p0, _ = net.predict(x)
    got = net.get_weights()
    assert len(got) == 108 and all(np.array_equal(a, b) for a, b in zip(got, weights_default))
    net.model.set_weights(weights_perturbed)                           # client.py:51
    p1, v1 = net.predict(x)
    pr, vr = nets["perturbed", "bf16"].predict(x)
    assert np.array_equal(p1, pr) and np.array_equal(v1, vr) and not np.array_equal(p0, p1)
    path = str(tmp_path / "w.npz")
    net.save(path)
    net.set_weights(weights_default)
    net.load(path)
    p2, _ = net.predict(x)
    assert np.array_equal(p2, p1)
    h5 = str(tmp_path / "iter_0.h5")                                  # the reference's checkpoint format (train.py:45)
    net.save(h5)
    net.set_weights(weights_default)
    net.load(h5)
    assert np.array_equal(net.predict(x)[0], p1)
    with pytest.raises(ValueError):
        net.set_weights(W.init_weights(blocks=2))
    net.close()


def test_predict_from_other_threads(nets):
    """usi.py:146-148 calls predict from a ponder thread; contexts serialise internally."""
    net = nets["default", "bf16"]
    x = synth.synthetic_planes(16, seed=8)
    want, _ = net.predict(x)
    out, errs = [None] * 4, []

    def work(i):
        try:
            for _ in range(5):
                out[i] = net.predict(x)[0]
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs and all(np.array_equal(o, want) for o in out)


def test_aggregation_service_matches_direct_predict(weights_default):
    """Many 'games' (threads) submit root (1) and leaf (16) batches; the service coalesces them into large GPU
    batches; every caller gets exactly what a direct predict of its own positions returns."""
    from erweitern_55_b200 import packing
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_default)
    games = 24
    xs = [synth.synthetic_planes(16, seed=900 + g) for g in range(games)]
    want = [net.predict(x) for x in xs]
    want1 = [net.predict(x[:1]) for x in xs]
    net.start_service(max_batch=256, max_wait_us=300)
    got, got1, gotp, errs = [None] * games, [None] * games, [None] * games, []

    def game(g):
        try:
            for _ in range(3):
                got1[g] = net.predict(xs[g][:1])                       # root evaluation, mcts.py:57-58
                got[g] = net.predict(xs[g])                            # leaf batch, mcts.py:101-102
                gotp[g] = net.predict_packed(*packing.pack_planes(xs[g]))
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=game, args=(g,)) for g in range(games)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    batches, positions = net.service_stats()
    net.stop_service()
    assert not errs
    assert positions == games * 3 * (1 + 16 + 16) and batches < games * 9      # requests really were coalesced
    for g in range(games):
        for a, b in ((got[g], want[g]), (got1[g], want1[g]), (gotp[g], want[g])):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    p, v = net.predict(xs[0])                                          # direct path works again after stop
    assert np.array_equal(p, want[0][0])
    m, e, mb = (net.start_service(1024, 200), net.bench_selfplay(workers=32, moves=1, simulations=64, leaf_batch=16))[1]
    assert m > 0 and e > 0 and mb > 1
    net.close()                                                        # destroy stops a running service


def test_native_selfplay_and_python_search_loop(weights_default):
    """The engine + MCTS + evaluation service end to end, and the reference's own search-loop shape in Python:
    positions -> get_inputs -> predict -> legal-move priors (mcts.py:91-106) with the native Position class."""
    from erweitern_55_b200.minishogi import Position
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_default)
    positions = []
    p = Position()
    rng = np.random.default_rng(3)
    for _ in range(16):
        moves = p.generate_moves()
        if not moves:
            break
        p.do_move(moves[int(rng.integers(len(moves)))])
        positions.append(p.copy(True))
    nninputs = net.get_inputs(positions)                                # network.py:260-274
    policy, value = net.predict(nninputs)
    value = (value + 1) / 2                                             # mcts.py:103
    bits = np.stack([q.to_packed_input()[0] for q in positions]); scale = np.stack([q.to_packed_input()[1] for q in positions])
    pp, vp = net.predict_packed(bits, scale)
    assert np.array_equal(pp, policy) and np.all((0 <= value) & (value <= 1))
    for b, q in enumerate(positions):
        idx = [q.policy_index(m) for m in q.generate_moves()]
        if idx:
            pri = np.exp(policy[b, idx] - policy[b, idx].max())
            assert np.isfinite(pri).all() and pri.sum() > 0
    net.start_service(max_batch=512, max_wait_us=200)
    st = net.selfplay(threads=16, total_moves=48, simulations=64, leaf_batch=16, mate_depth=3, seed=7)
    net.stop_service()
    assert st["moves"] >= 40 and st["evals"] > st["moves"] * 40 and st["mean_gpu_batch"] > 4
    net.close()


def test_legal_move_priors_match_the_host_softmax(weights_perturbed):
    """e55_predict_packed_moves: softmax over the legal moves' logits on the GPU == the same softmax of the logits
    predict_packed returns, directly and through the evaluation service; ragged and empty move lists included."""
    from erweitern_55_b200.minishogi import Position, encode_packed_batch
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_perturbed)
    rng = np.random.default_rng(9)
    positions, p = [], Position()
    while len(positions) < 40:
        moves = p.generate_moves()
        if not moves or p.get_ply() > 30:
            p = Position()
            continue
        p.do_move(moves[int(rng.integers(len(moves)))])
        positions.append(p.copy(True))
    lists = [[q.policy_index(m) for m in q.generate_moves()] for q in positions]
    lists[3] = []                                                       # a position without moves gets no priors
    lists[7] = list(range(0, 1725, 7))                                  # a long list (247 entries)
    off = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int32)
    idx = np.array([i for l in lists for i in l], dtype=np.uint16)
    bits, scale = encode_packed_batch(positions)
    logits, value = net.predict_packed(bits, scale)

    def check(priors, v):
        assert np.array_equal(v, value)
        for b, l in enumerate(lists):
            got = priors[off[b]:off[b + 1]]
            if not l:
                continue
            e = np.exp(logits[b, l].astype(np.float64) - logits[b, l].max())
            assert np.abs(got - e / e.sum()).max() < 2e-6 and abs(got.sum() - 1) < 1e-5
        # the tower kernel's fused tail == the stand-alone softmax kernel on the same logits, bit for bit
        alone = np.empty_like(priors)
        net._check(net._lib.e55_debug_legal_priors(net._ctx, logits.ctypes.data, len(lists), off.ctypes.data, idx.ctypes.data, alone.ctypes.data))
        assert np.array_equal(alone, priors)

    check(*net.predict_packed_moves(bits, scale, off, idx))
    net.set_precision("fp32")
    lf, vf = net.predict_packed(bits, scale)
    pf, _ = net.predict_packed_moves(bits, scale, off, idx)
    e = np.exp(lf[0, lists[0]] - lf[0, lists[0]].max())
    assert np.abs(pf[:len(lists[0])] - e / e.sum()).max() < 2e-6
    net.set_precision("bf16")
    net.start_service(max_batch=64, max_wait_us=100)
    out, errs = {}, []

    def worker(lo, hi):
        try:
            o = off[lo:hi + 1] - off[lo]
            out[lo] = net.predict_packed_moves(bits[lo:hi], scale[lo:hi], o, idx[off[lo]:off[hi]])
        except Exception as ex:  # pragma: no cover
            errs.append(ex)
    ts = [threading.Thread(target=worker, args=(lo, min(lo + 6, 40))) for lo in range(0, 40, 6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    net.stop_service()
    assert not errs
    check(np.concatenate([out[lo][0] for lo in sorted(out)]), np.concatenate([out[lo][1] for lo in sorted(out)]))
    with pytest.raises(Exception):
        net.predict_packed_moves(bits[:1], scale[:1], np.array([0, 1], np.int32), np.array([1725], np.uint16))
    # asynchronous form: several tickets in flight from ONE thread, fetched out of order
    import ctypes as C
    from erweitern_55_b200 import _capi
    lib = _capi.load()
    net.start_service(max_batch=16, max_wait_us=100)
    tickets, spans = [], [(0, 10), (10, 16), (16, 33), (33, 40)]          # 17 > what is left of the first batch of 16
    for lo, hi in spans:
        t = (C.c_uint8 * 32)()
        o = np.ascontiguousarray(off[lo:hi + 1] - off[lo])
        rc = lib.e55_service_submit_packed_moves(net._ctx, bits[lo:hi].ctypes.data, scale[lo:hi].ctypes.data, hi - lo,
                                                 o.ctypes.data, idx[off[lo]:off[hi]].ctypes.data, 1, t)
        assert rc == (_capi.E55_OK if hi - lo <= 16 else _capi.E55_ERR_INVALID)
        tickets.append(t if rc == 0 else None)
    got = {}
    for k in (3, 0, 1):
        lo, hi = spans[k]
        pr = np.empty(off[hi] - off[lo], np.float32); va = np.empty((hi - lo, 1), np.float32)
        assert lib.e55_service_fetch(net._ctx, tickets[k], pr.ctypes.data, va.ctypes.data, 1) == 0
        got[k] = (pr, va)
    net.stop_service()
    ref_p, ref_v = net.predict_packed_moves(bits, scale, off, idx)
    for k in (0, 1, 3):
        lo, hi = spans[k]
        assert np.array_equal(got[k][0], ref_p[off[lo]:off[hi]]) and np.array_equal(got[k][1], ref_v[lo:hi])
    net.close()


class _ReferenceStyleNetwork:
    """Only the reference's three calls (get_input / get_inputs / predict), so mcts.MCTS.run takes its leaf-by-leaf
    Python path (mcts.py:91-110) instead of the native batch step."""

    def __init__(self, net):
        self.get_input, self.get_inputs, self.predict = net.get_input, net.get_inputs, net.predict


def test_native_search_equals_the_reference_style_python_loop(weights_perturbed):
    """mcts.MCTS.run on the native batch step (bit-packed inputs, one C call per leaf batch) builds exactly the tree
    the reference's Python loop builds with the same Network: same visit counts, same principal variation."""
    from erweitern_55_b200 import mcts
    from erweitern_55_b200.minishogi import Position
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_perturbed)
    results = []
    for nn in (net, _ReferenceStyleNetwork(net)):
        cfg = mcts.Config()
        cfg.simulation_num = 256
        search = mcts.MCTS(cfg)
        p = Position()
        dumps = []
        for _ in range(3):                                              # three moves: tree reuse included
            root = search.run(p, nn)
            dumps.append((search.dump(root), [m.sfen() for m in search.mcts.info(root)[0]]))
            p.do_move(search.best_move(root))
        results.append(dumps)
    # the native step takes its priors from the GPU's legal-move softmax (__expf), the Python loop from the host's
    # exp: the same search up to last-bit prior differences, which can only flip near-ties
    for (dn, pvn), (dp, pvp) in zip(*results):
        assert dn[0] == dp[0] and abs(dn[1] - dp[1]) < 0.02
        vn, vp = dict(dn[2]), dict(dp[2])
        tv = sum(abs(vn.get(m, 0) - vp.get(m, 0)) for m in set(vn) | set(vp)) / (2 * dn[0])
        assert tv < 0.1, tv
    assert results[0][0][0][0] == 256 and results[0][1][0][0] >= 256    # reused subtree adds to the next search
    pp, vp = net.predict_positions([p])
    pf, vf = net.predict(net.get_inputs([p]))
    assert np.array_equal(pp, pf) and np.array_equal(vp, vf)
    net.close()


def test_selfplay_records_and_usi_on_the_gpu(weights_default):
    import io
    from erweitern_55_b200 import mcts, selfplay, usi
    from erweitern_55_b200.minishogi import Move, Position
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_default)
    # one game through the reference-shaped selfplay.run (native batch step underneath)
    cfg = mcts.Config()
    cfg.simulation_num = 64
    cfg.use_dirichlet = True
    rec = selfplay.run(net, mcts.MCTS(cfg), max_moves=12)
    assert rec.ply == len(rec.sfen_kif) == len(rec.mcts_result) > 4
    # many games on the native driver, records through the sink
    net.start_service(max_batch=512, max_wait_us=200)
    stats, games = selfplay.run_many(net, threads=4, games_per_thread=8, total_moves=400, simulations=32, leaf_batch=16,
                                     mate_depth=3, max_moves=20, seed=11)
    auto, _ = selfplay.run_many(net, total_moves=200, simulations=32, max_moves=20, records=False)   # thread counts from the cores
    net.stop_service()
    assert auto["moves"] >= 150
    assert stats["games"] == len(games) >= 8 and stats["mean_gpu_batch"] > 4
    for g in games:
        q = Position()
        for move, (sum_n, qv, visits) in zip(g.sfen_kif, g.mcts_result):
            assert move in {m.sfen() for m in q.generate_moves()} and sum_n == sum(n for _, n in visits)
            q.do_move(Move(move))
    # USI in latency mode on the direct (non-service) path
    out = io.StringIO()
    engine = usi.USI(nn=net, stdout=out, stdin=io.StringIO(
        "usi\nisready\nposition startpos moves 5d5c\ngo btime 2000 wtime 2000 byoyomi 0\nquit\n"))
    engine.start()
    lines = out.getvalue().splitlines()
    best = [l.split()[1] for l in lines if l.startswith("bestmove")]
    start = Position(); start.do_move(Move("5d5c"))
    assert len(best) == 1 and best[0] in {m.sfen() for m in start.generate_moves()}
    nodes = [int(l.split()[4]) for l in lines if l.startswith("info depth")]
    assert nodes and nodes[-1] > 2000                                   # 200 ms of batch-16 forwards at ~100 us each
    net.close()


def test_host_side_packing_is_invisible(weights_perturbed):
    """e55_predict bit-packs large float-plane batches on a host thread pool before the copy (e55_host.cpp): outputs are
    bit-identical to the plain float path, also when some chunk cannot be packed exactly and travels as floats."""
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_perturbed, max_batch=1024)
    lib = net._lib
    x = synth.synthetic_planes(1000, seed=77)                          # ragged: not a multiple of the chunk
    lib.e55_set_host_threads(net._ctx, 0)
    p0, v0 = net.predict(x)
    import ctypes as C
    from erweitern_55_b200 import hostmem
    xp = hostmem.empty(x.shape)                                        # page-locked copy of the same planes: the copy engine
    xp[...] = x                                                        # may take a share of them as floats
    for planes, pinned in ((x, False), (xp, True)):
        for threads in (1, 4, -1):
            lib.e55_set_host_threads(net._ctx, threads)
            shares = set()
            for _ in range(18 if threads < 0 else 2):                  # the automatic mode tries every split of the batch
                p1, v1 = net.predict(planes)                           # between packed chunks and float chunks first
                assert np.array_equal(p0, p1) and np.array_equal(v0, v1)
                eighths = C.c_int()
                assert lib.e55_get_host_split(net._ctx, C.byref(eighths), None) == 0
                shares.add(eighths.value)
            assert shares == ({0, 1, 2, 3, 4, 5, 6, 8} if threads < 0 and pinned else {0})
    # results land in the caller's arrays whatever memory those are: page-locked (Network.predict's own), pageable numpy
    lib.e55_set_host_threads(net._ctx, -1)
    for n in (64, 100, 300, 511, 1000):
        for planes in (x[:n], xp[:n]):
            pol, val = np.full((n, 1725), np.nan, dtype=np.float32), np.full((n, 1), np.nan, dtype=np.float32)
            net._check(lib.e55_predict(net._ctx, planes.ctypes.data, n, pol.ctypes.data, val.ctypes.data))
            assert np.array_equal(pol, p0[:n]) and np.array_equal(val, v0[:n])
    y = x.copy()
    y[300, 5, 2, 2], y[300, 5, 3, 3] = 0.25, 0.75                      # two different non-zero values in one plane
    y[999, 265] = np.linspace(0.0, 1.0, 25, dtype=np.float32).reshape(5, 5)
    y[5, 7, 0, 0] = -0.0                                               # negative zero is still zero
    lib.e55_set_host_threads(net._ctx, 0)
    q0, w0 = net.predict(y)
    for threads, calls in ((4, 1), (-1, 12)):
        lib.e55_set_host_threads(net._ctx, threads)
        for _ in range(calls):
            q1, w1 = net.predict(y)
            assert np.array_equal(q0, q1) and np.array_equal(w0, w1) and not np.array_equal(q0[300], p0[300])
    net.close()


def test_device_history_and_encoding_equal_the_hosts():
    """What the network sees in the GPU-resident self-play: random games plus random descents below them, played with the
    kernels' own bookkeeping (game history ring, path view, Bloom-filtered repetition counts) and encoded on the device,
    against the host Position replaying the same moves -- identical 266 packed planes, repetition verdicts and counts,
    legal-move counts."""
    import ctypes as C
    from erweitern_55_b200 import _capi
    lib = _capi.load()
    bad, reps, path = C.c_int64(-1), C.c_int64(), C.c_int64()
    assert lib.e55_selftest_encode(0, 8192, 2024, C.byref(bad), C.byref(reps), C.byref(path)) == 0
    assert bad.value == 0, f"{bad.value} of 8192 positions differ"
    assert reps.value > 500 and path.value > 8192 * 10           # repetitions and real descents were exercised


@pytest.mark.parametrize("cohorts", ["1", "2"])
def test_gpu_resident_selfplay(weights_default, monkeypatch, cohorts):
    """The search on the GPU (csrc/e55_gpu_selfplay.cuh): the device build of the rules engine counts the same perft
    nodes as the host build, and every game it plays is a legal game that ended for a reason the rules give -- with
    all games in one cohort and with two cohorts on two streams (the layout large runs use)."""
    monkeypatch.setenv("E55_GPU_COHORTS", cohorts)
    import ctypes as C
    from erweitern_55_b200 import _capi
    from erweitern_55_b200.minishogi import Move, Position
    from erweitern_55_b200.network import Network
    lib = _capi.load()
    for depth, nodes in ((1, 14), (2, 181), (3, 2512), (4, 35401)):
        got = C.c_int64()
        assert lib.e55_selftest_engine(0, depth, C.byref(got)) == 0 and got.value == nodes
    net = Network(precision="bf16", weights=weights_default)
    stats, games = net.gpu_selfplay(games=512, total_moves=40000, simulations=32, records=256, seed=9, max_moves=80)
    assert stats["tree_overflows"] == 0 and stats["records_truncated"] == 0 and stats["mate_searches"] > 0
    assert all(g.complete for g in games)
    assert stats["moves"] >= 40000 and stats["games"] >= 30 and len(games) >= 30
    assert 32 * 0.3 * stats["moves"] < stats["evals"] <= 33 * (stats["moves"] + 512)   # one root + <= 32 leaves per searched move
    reasons = {"no moves": 0, "repetition": 0, "length": 0}
    mates = 0
    for rec in games:                                                   # GameRecord objects, as selfplay.run returns them
        p = Position()
        assert rec.complete and rec.ply == len(rec.sfen_kif) == len(rec.mcts_result) <= 80
        for m, (sum_n, q, visits) in zip(rec.sfen_kif, rec.mcts_result):
            legal = {x.sfen() for x in p.generate_moves()}
            assert m in legal and {v for v, _ in visits} <= legal and 0.0 <= q <= 1.0, (m, p.sfen())
            assert sum_n == sum(n for _, n in visits) and dict(visits).get(m, 0) > 0
            if (sum_n, visits) == (1, [(m, 1)]):                         # a mate-search move (selfplay.py:82-84)
                mates += 1
                assert p.solve_checkmate_dfs(7)[0]
            else:
                assert sum_n >= 32                                       # a full search (more when the tree was reused)
            p.do_move(Move(m))
        rep, my_check, op_check = p.is_repetition()
        if rep:
            reasons["repetition"] += 1
            assert rec.winner == (1 - p.get_side_to_move() if my_check else p.get_side_to_move() if op_check else 1)
        elif not p.generate_moves():
            reasons["no moves"] += 1
            assert rec.winner == 1 - p.get_side_to_move()
        else:
            reasons["length"] += 1
            assert rec.ply == 80 and rec.winner == 2
    assert reasons["no moves"] > 0 and mates > 0
    # without mate search and tree reuse (the options of selfplay.run / mcts.Config) games are played just as legally
    plain, games2 = net.gpu_selfplay(games=256, total_moves=12000, simulations=32, mate_depth=0, reuse_tree=False, records=64, seed=4, max_moves=60)
    assert plain["moves"] >= 12000 and len(games2) >= 5
    for rec in games2:
        p = Position()
        for m, (sum_n, q, visits) in zip(rec.sfen_kif, rec.mcts_result):
            assert m in {x.sfen() for x in p.generate_moves()} and 32 <= sum_n <= 56, (m, p.sfen(), sum_n)
            p.do_move(Move(m))
    with pytest.raises(Exception):
        net.gpu_selfplay(games=64, total_moves=10, simulations=100)                     # not a multiple of 16
    net.close()


def test_device_resident_api(nets):
    net = nets["default", "bf16"]
    x = synth.synthetic_planes(100, seed=12)
    p, v = net.predict(x)
    xd = torch.from_numpy(x).cuda()
    pd = torch.empty((100, 1725), device="cuda")
    vd = torch.empty((100, 1), device="cuda")
    before = net.launch_count()
    net.predict_device(xd, pd, vd)
    net.sync()
    assert net.launch_count() > before
    assert np.array_equal(pd.cpu().numpy(), p) and np.array_equal(vd.cpu().numpy(), v)


def test_small_and_scaled_architectures(weights_default):
    """Other tower shapes on both paths: 2 blocks x 128 filters on 40 input planes, 2 blocks x 256 filters
    (the width of BASELINE.json configs[4]; the tcgen05 kernel has a 256-filter instantiation)."""
    from erweitern_55_b200.network import Network
    O = _oracle()
    w2 = W.init_weights(blocks=2, filters=128, in_planes=40, seed=5, perturbed=True)
    x = synth.synthetic_planes(9, seed=2, in_planes=40)
    po, vo = O.forward_torch(w2, x)
    for prec, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL_PERTURBED)):
        net = Network(blocks=2, in_planes=40, precision=prec, weights=w2)
        _check(*net.predict(x), po, vo, tol)
        net.close()
    w3 = W.init_weights(blocks=2, filters=256, seed=6, perturbed=True)
    for n in (1, 5, 37):
        x = synth.synthetic_planes(n, seed=3 + n)
        po, vo = O.forward_torch(w3, x)
        for prec, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL_PERTURBED)):
            net = Network(blocks=2, filters=256, precision=prec, weights=w3)
            _check(*net.predict(x), po, vo, tol)
            if prec == "bf16":
                a = net.debug_activations(x, 5)
                net.set_precision("fp32")
                b = net.debug_activations(x, 5)
                assert np.abs(a - b).max() <= 0.02 * max(1.0, np.abs(b).max())
            net.close()


def test_scaled_tower_20x256():
    """BASELINE.json configs[4]: 20 residual blocks x 256 filters.  With Keras-default init (identity BN) the
    activations of a 43-layer tower grow, so the bf16 tolerance is stated relative to the logit scale:
    |dlogit| <= 1 % of max|logit| (the 7x128 net's 0.03 is 1.5 % of its +-2 range), |dvalue| <= 0.03."""
    from erweitern_55_b200.network import Network
    O = _oracle()
    w = W.init_weights(blocks=20, filters=256, seed=55)
    x = synth.synthetic_planes(24, seed=31)
    po, vo = O.forward_torch(w, x)
    net = Network(blocks=20, filters=256, precision="bf16", weights=w)
    p, v = net.predict(x)
    scale = float(np.abs(po).max())
    assert np.abs(p - po).max() <= 0.01 * scale, (np.abs(p - po).max(), scale)
    assert np.abs(v - vo).max() <= 0.03
    top2 = np.sort(po, axis=1)[:, -2:]
    decisive = (top2[:, 1] - top2[:, 0]) > 0.02 * scale
    assert (p.argmax(1)[decisive] == po.argmax(1)[decisive]).all()
    net.close()


def test_umma_addressing_selftest():
    import ctypes as C
    from erweitern_55_b200 import _capi
    lib = _capi.load()
    for shift in (0, 1, 7, 25, 49):
        err = C.c_float(-1)
        assert lib.e55_selftest_umma(0, shift, C.byref(err)) == 0
        assert err.value == 0.0


@pytest.mark.parametrize("prefix,simulations,searches,full", [
    ([], 800, 3, False),                                              # the opening, as self-play searches it; tree reuse twice
    (["5e4d", "1a2b", "4d5e", "2b1a", "5e4d", "1a2b", "4d5e", "2b1a", "5e4d", "1a2b"], 480, 2, False),   # a repetition within reach of the search
    ([], 640, 2, True),                                               # "full" searches of playout-cap oscillation (selfplay.py:47-53):
])                                                                    # forced playouts, fresh tree per move, pruned targets
def test_gpu_search_walks_like_the_oracle(weights_default, prefix, simulations, searches, full):
    """select_kernel / expand_kernel (GPU-resident self-play) against oracle/mcts_oracle.py, selection by selection, through
    the debug stepper of the C ABI: tests/search_walk.py (the same walk runs on the host build of the kernels in the CPU
    suite, tests/test_search_on_host.py)."""
    from erweitern_55_b200.network import Network
    from search_walk import CudaStepper, walk_like_the_oracle
    net = Network(precision="bf16", weights=weights_default)
    _, terminals, _, _ = walk_like_the_oracle(CudaStepper(net), prefix, simulations, searches, full)
    net.close()
    if prefix:
        assert terminals > 0                                          # the walk did run into fourth occurrences


def test_channel_split_kernel_is_bit_identical(weights_perturbed):
    """csrc/e55_tc_split.cuh (an experiment that did not pay and is off by default: E55_SPLIT_MAX switches it on): the two
    CTAs of a cluster split every layer's output channels and exchange activations through distributed shared memory.
    Its results must equal the throughput kernel's bit for bit -- logits, values, priors of the fused tail -- and its
    trunk the fp32 path's within the bf16 tolerance."""
    import os
    from erweitern_55_b200 import packing
    from erweitern_55_b200.network import Network
    ref = Network(precision="bf16", weights=weights_perturbed)
    os.environ["E55_SPLIT_MAX"] = "64"
    try:
        net = Network(precision="bf16", weights=weights_perturbed)
    finally:
        del os.environ["E55_SPLIT_MAX"]
    rng = np.random.default_rng(5)
    for n in (1, 3, 4, 5, 16, 37, 64):
        x = synth.synthetic_planes(n, seed=700 + n)
        p0, v0 = ref.predict(x)
        p1, v1 = net.predict(x)
        assert np.array_equal(p0, p1) and np.array_equal(v0, v1), n
        bits, scale = packing.pack_planes(x)
        counts = rng.integers(0, 50, size=n)
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        idx = rng.integers(0, 1725, size=int(off[-1])).astype(np.uint16)
        q0, w0 = ref.predict_packed_moves(bits, scale, off, idx)
        q1, w1 = net.predict_packed_moves(bits, scale, off, idx)
        assert np.array_equal(q0, q1) and np.array_equal(w0, w1), n
    x = synth.synthetic_planes(9, seed=3)
    assert np.array_equal(ref.debug_activations(x, 5), net.debug_activations(x, 5))
    ref.close(); net.close()


def test_playout_cap_oscillation_in_the_native_drivers(weights_default):
    """selfplay.py:45-61,90-91 in the host driver (e55_selfplay_run_ex) and in the GPU-resident one: fast searches (n
    simulations, reused tree) are no learning targets, full searches (N, forced playouts, fresh tree, pruned counts)
    and mate moves are; the games stay legal."""
    from erweitern_55_b200 import selfplay
    from erweitern_55_b200.minishogi import Move, Position
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_default)
    cap = {"enable": True, "N": 96, "n": 32, "frac": 0.4}

    def check(games):
        targets = plies = full_sum = 0
        for rec in games:
            p = Position()
            for ply, (m, (sum_n, q, visits)) in enumerate(zip(rec.sfen_kif, rec.mcts_result)):
                legal = {x.sfen() for x in p.generate_moves()}
                assert m in legal and all(v in legal for v, _ in visits)
                is_target = ply in rec.learning_target_plys
                is_mate = (sum_n, visits) == (1, [(m, 1)])
                if not is_mate:
                    if is_target:
                        assert sum_n <= cap["N"] + 24 and sum(n for _, n in visits) == sum_n   # pruned: only what is kept is counted
                        full_sum += sum_n
                    else:
                        assert sum_n >= cap["n"]                                            # the reused tree's visits count
                targets += is_target; plies += 1
                p.do_move(Move(m))
        assert 0.2 * plies < targets < 0.8 * plies, (targets, plies)
        return targets, plies

    stats, games = net.gpu_selfplay(games=256, total_moves=6000, simulations=96, records=128, seed=3, max_moves=60, mate_depth=3,
                                    playout_cap_oscillation=cap)
    assert stats["tree_overflows"] == 0 and len(games) >= 32
    check(games)
    net.start_service(max_batch=256, max_wait_us=200)
    stats, games = selfplay.run_many(net, threads=2, total_moves=1500, simulations=96, leaf_batch=16, mate_depth=3, seed=5, max_moves=60,
                                     playout_cap_oscillation=cap)
    net.stop_service()
    assert len(games) >= 10
    check(games)
    net.close()


def test_layout_adapter_on_the_gpu(weights_perturbed):
    """Network.set_layout_adapter folds a foreign plane / policy-channel layout into the installed weights: same results
    as installing the adapted list, get_weights() still the checkpoint's own list, and no effect once it is cleared."""
    from erweitern_55_b200 import weights as W
    from erweitern_55_b200.network import Network
    rng = np.random.default_rng(2)
    plane_perm, policy_perm = rng.permutation(266), rng.permutation(69)
    plane_scale = rng.choice([1.0, 0.5], size=266).astype(np.float32)
    x = synth.synthetic_planes(9, seed=4)
    a = Network(precision="bf16", weights=weights_perturbed)
    p_plain, v_plain = a.predict(x)
    a.set_layout_adapter(plane_perm, plane_scale, policy_perm)
    b = Network(precision="bf16", weights=W.adapt_weights(weights_perturbed, plane_perm, plane_scale, policy_perm))
    pa, va = a.predict(x)
    pb, vb = b.predict(x)
    assert np.array_equal(pa, pb) and np.array_equal(va, vb) and not np.array_equal(pa, p_plain)
    assert all(np.array_equal(u, v) for u, v in zip(a.get_weights(), weights_perturbed))
    a.set_layout_adapter()
    p2, v2 = a.predict(x)
    assert np.array_equal(p2, p_plain) and np.array_equal(v2, v_plain)
    a.close(); b.close()


# ---- launches far beyond the grid (last in the file: whatever happens here, every other test has run) --------------------
@pytest.mark.timeout(300, method="thread")   # a launch that never ends must end the run, not hang it
def test_one_launch_of_many_rounds(weights_perturbed):
    """A single launch far beyond the grid (120 rounds per CTA: what a GPU self-play round of 8192 games x 16 leaves is) gives
    every position the result it has in a one-round launch -- per-launch state that only shows after hundreds of input
    slices per CTA (barrier phases, slice counters) has nowhere to hide."""
    from erweitern_55_b200.network import Network
    reps, nb = 120, 1184
    net = Network(precision="bf16", weights=weights_perturbed, max_batch=nb)
    xb = torch.from_numpy(synth.synthetic_planes(nb, seed=31)).cuda()
    pb = torch.empty((nb, 1725), device="cuda")
    vb = torch.empty((nb, 1), device="cuda")
    net.predict_device(xb, pb, vb)
    net.sync()
    x = xb.repeat(reps, 1, 1, 1).contiguous()
    p = torch.full((reps * nb, 1725), float("nan"), device="cuda")
    v = torch.full((reps * nb, 1), float("nan"), device="cuda")
    net.predict_device(x, p, v)
    net.sync()
    assert torch.equal(p.view(reps, nb, 1725), pb.unsqueeze(0).expand(reps, nb, 1725))
    assert torch.equal(v.view(reps, nb, 1), vb.unsqueeze(0).expand(reps, nb, 1))
    n_odd = 100 * nb + 5                                               # ragged: the last round is partly empty
    p.fill_(float("nan")); v.fill_(float("nan"))
    net.predict_device(x[:n_odd], p[:n_odd], v[:n_odd])
    net.sync()
    assert torch.equal(p[:n_odd], pb.repeat(101, 1)[:n_odd]) and torch.equal(v[:n_odd], vb.repeat(101, 1)[:n_odd])
    assert bool(torch.isnan(p[n_odd:]).all()) and bool(torch.isnan(v[n_odd:]).all())
    net.close()
    del x, p, v
    torch.cuda.empty_cache()


@pytest.mark.timeout(300, method="thread")   # a launch that never ends must end the run, not hang it
def test_gpu_selfplay_at_the_bench_size(weights_default):
    """16384 games in flight (two cohorts, 131072 network slots per launch on 124 SMs): one move per game, no overflow."""
    from erweitern_55_b200.network import Network
    net = Network(precision="bf16", weights=weights_default)
    stats, _ = net.gpu_selfplay(games=16384, total_moves=16384, simulations=32, seed=3)
    assert stats["moves"] >= 16384 and stats["tree_overflows"] == 0 and stats["evals"] >= 16384 * 16
    net.close()
