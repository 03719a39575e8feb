"""Host-side mirror of the reference interface: weight list format, input stacking, sharding."""
import numpy as np
import pytest

from erweitern_55_b200 import sharding, synth, weights as W
from erweitern_55_b200.network import Network, _parse_device


def test_weight_list_matches_reference_architecture():
    spec = W.weight_spec()
    assert len(spec) == 108                       # 18 convs x2 + 17 BN x4 + 2 dense x2
    assert W.num_params() == 2_541_003
    assert spec[0] == ("stem_conv/kernel", (3, 3, 266, 128))
    w = W.init_weights(seed=1)
    assert [t.shape for t in w] == [s for _, s in spec]
    assert all(t.dtype == np.float32 for t in w)
    assert W.infer_arch(w) == (7, 128, 266)
    assert W.infer_arch(W.init_weights(blocks=20, filters=256, seed=1)) == (20, 256, 266)
    with pytest.raises(ValueError):
        W.infer_arch(w[:-1])


def test_default_init_is_keras_default():
    w = dict(zip([n for n, _ in W.weight_spec()], W.init_weights(seed=3)))
    assert np.all(w["stem_conv/bias"] == 0) and np.all(w["stem_bn/gamma"] == 1)
    assert np.all(w["stem_bn/moving_mean"] == 0) and np.all(w["stem_bn/moving_variance"] == 1)
    lim = np.sqrt(6.0 / (9 * 266 + 9 * 128))
    k = w["stem_conv/kernel"]
    assert np.abs(k).max() <= lim and np.abs(k).max() > 0.95 * lim


def test_synthetic_planes_contract():
    x = synth.synthetic_planes(5, seed=2)
    assert x.shape == (5, 266, 5, 5) and x.dtype == np.float32
    assert x.min() >= 0.0 and x.max() <= 1.0            # network.py:169-170
    assert np.all(x[:, 0] == x[:, 0, :1, :1]) and np.all(x[:, 1] == x[:, 1, :1, :1])
    m = synth.synthetic_legal_mask(4)
    assert m.shape == (4, 1725) and np.all(m.sum(1) == 30)


class _FakePosition:
    def __init__(self, fill):
        self.fill = fill

    def to_alphazero_input(self):
        return [self.fill] * (266 * 25)


def test_input_stacking_matches_reference_semantics():
    """network.py:255-274,283-286 without constructing a device context."""
    net = Network.__new__(Network)
    net.input_shape = [266, 5, 5]
    a = net.get_input(_FakePosition(0.5))
    assert a.shape == (266, 5, 5)
    assert net.get_input(_FakePosition(0.5), True).shape == (1, 266, 5, 5)
    ins = net.get_inputs([_FakePosition(0.25), _FakePosition(1.0)])
    assert ins.shape == (2, 266, 5, 5) and ins.dtype == np.float32
    assert np.all(ins[0] == 0.25) and np.all(ins[1] == 1.0)
    z = net.zero_inputs(3)
    assert z.shape == (3, 266, 5, 5) and z.dtype == np.float32 and not z.any()
    with pytest.raises(NotImplementedError):
        net.step(None, None, None)


def test_device_parsing():
    assert _parse_device("gpu") == 0 and _parse_device("cuda:3") == 3 and _parse_device(2) == 2
    for bad in ("cpu", "tpu"):
        with pytest.raises(ValueError):
            _parse_device(bad)


def test_shard_ranges_partition_the_batch():
    for total in (1, 7, 1024, 1000):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_pack_planes_roundtrip_and_rejection():
    from erweitern_55_b200 import packing
    x = synth.synthetic_planes(7, seed=9)
    bits, scale = packing.pack_planes(x)
    assert bits.shape == (7, 266) and bits.dtype == np.uint32 and scale.shape == (7, 266) and scale.dtype == np.float32
    assert bits.max() < (1 << 25)
    np.testing.assert_array_equal(packing.unpack_planes(bits, scale), x)
    assert np.all(bits[:, 0][x[:, 0, 0, 0] == 1] == (1 << 25) - 1)          # constant plane: every square set
    assert np.array_equal(scale[:, 1], x[:, 1, 0, 0])                        # ply plane keeps its value
    bad = x.copy()
    bad[0, 5, 0, 0], bad[0, 5, 0, 1] = 0.25, 0.5
    with pytest.raises(ValueError):
        packing.pack_planes(bad)
    with pytest.raises(ValueError):
        packing.pack_planes(np.zeros((2, 3, 4, 5)))


def test_result_array_pool_without_a_device():
    """hostmem.empty is np.empty to its callers: right shape and dtype whatever the size, views keep their block alive,
    and without a CUDA device (no page-locked memory to be had) it hands out ordinary arrays instead of failing."""
    import gc
    import torch
    from erweitern_55_b200 import hostmem
    for shape, dtype in (((16, 1725), np.float32), ((1024, 1725), np.float32), ([3, 266, 5, 5], np.float32), ((70000,), np.uint8)):
        a = hostmem.empty(shape, dtype)
        assert isinstance(a, np.ndarray) and a.shape == tuple(shape) and a.dtype == np.dtype(dtype) and a.flags.c_contiguous and a.flags.writeable
        a[...] = 1
        v = a[1:] if a.ndim > 1 else a[10:]
        del a
        gc.collect()
        assert (v == 1).all()                                          # the view holds the memory
    if not torch.cuda.is_available():
        assert hostmem.empty((1024, 1725)).base is None                # pageable: nothing page-locked without a device
    hostmem.trim()
