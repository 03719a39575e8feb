"""The minimal HDF5 reader/writer behind Network.load/save of Keras weight files.

The reader is pinned on a file written by the real HDF5 library: scipy ships a MATLAB v7.3 test file (HDF5 with a
512-byte user block; old-style groups, object headers v1, attributes -- the structures h5py-written Keras files use).
The writer is checked through the reader and, message by message, against the encodings in that file."""
import os

import numpy as np
import pytest

from erweitern_55_b200 import hdf5, weights as W

REAL = os.path.join(os.path.dirname(__import__("scipy").__file__), "io", "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
needs_real = pytest.mark.skipif(not os.path.exists(REAL), reason="scipy's HDF5 sample file is not installed")


@needs_real
def test_reader_on_a_file_written_by_libhdf5():
    f = hdf5.File(REAL)
    assert f.base == 512 and f.keys() == ["testdouble"] and "testdouble" in f and "nope" not in f
    d = f["testdouble"]
    assert isinstance(d, hdf5.Dataset) and d.shape == (9, 1) and d.dtype == np.float64
    assert np.allclose(d.read().ravel(), np.arange(0, 2 * np.pi + 1e-9, np.pi / 4))      # what scipy's own test expects
    assert d.attrs == {"MATLAB_class": b"double"}
    with pytest.raises(KeyError):
        f["testdouble/deeper"]


@needs_real
def test_writer_encodes_messages_like_libhdf5():
    d = hdf5.File(REAL)["testdouble"]
    real_dt, real_ds = d.first(0x0003), d.first(0x0001)
    mine_dt = hdf5.Writer._datatype(np.zeros(1, np.float64))
    assert real_dt[:len(mine_dt)] == mine_dt and not any(real_dt[len(mine_dt):])        # IEEE double, little endian
    assert real_ds == hdf5.Writer._dataspace((9, 1))


def test_round_trip_of_groups_datasets_and_attributes():
    w = hdf5.Writer()
    w.attrs["title"] = np.bytes_(b"erweitern")
    w.attrs["numbers"] = np.arange(5, dtype=np.int32)
    rng = np.random.default_rng(0)
    data = {f"layer_{i:02d}/w{i}:0": rng.standard_normal((i + 1, 3)).astype(np.float32) for i in range(40)}   # > one symbol node
    for k, v in data.items():
        w.create_dataset(k, v)
    w.create_dataset("scalars/f64", np.array([1.5, -2.5]))
    w.require_group("empty").attrs["weight_names"] = np.zeros((0,), "S1")
    f = hdf5.File(w.tobytes())
    assert f.attrs["title"] == b"erweitern" and list(f.attrs["numbers"]) == [0, 1, 2, 3, 4]
    assert sorted(f.keys()) == sorted({k.split("/")[0] for k in data} | {"scalars", "empty"})
    for k, v in data.items():
        got = f[k].read()
        assert got.dtype == np.float32 and np.array_equal(got, v)
    assert np.array_equal(f["scalars/f64"].read(), [1.5, -2.5]) and f["empty"].keys() == []
    with pytest.raises(ValueError):
        hdf5.File(b"not an hdf5 file at all" * 40)


def test_keras_weight_file_round_trip(tmp_path):
    ws = W.init_weights(blocks=2, filters=128, seed=3, perturbed=True)
    layers = W.keras_layers(blocks=2, filters=128)
    assert [n for n, _ in layers][:4] == ["conv2d", "batch_normalization", "conv2d_1", "batch_normalization_1"]
    assert layers[0][1] == ["conv2d/kernel:0", "conv2d/bias:0"] and layers[-1][0] == "dense_1"
    assert sum(len(names) for _, names in layers) == len(ws)
    path = tmp_path / "iter_0.h5"
    hdf5.write_keras_weights(path, ws, layers)
    back = hdf5.read_keras_weights(path)
    assert len(back) == len(ws) and all(np.array_equal(a, b) for a, b in zip(back, ws))
    f = hdf5.File(path)
    assert f.attrs["backend"] == b"tensorflow" and f["conv2d_1"]["conv2d_1"]["kernel:0"].shape == (3, 3, 128, 128)
    # the layout keras.Model.save uses (network.py:245-253): the same tree under /model_weights, other groups beside it
    full = hdf5.Writer()
    mw = full.require_group("model_weights")
    mw.attrs["layer_names"] = np.array([n.encode() for n, _ in layers] + [b"activation"])     # a layer without weights
    i = 0
    for name, names in layers:
        g = mw.require_group(name)
        g.attrs["weight_names"] = np.array([n.encode() for n in names])
        for n in names:
            g.create_dataset(n, ws[i])
            i += 1
    mw.require_group("activation").attrs["weight_names"] = np.zeros((0,), "S1")
    full.require_group("optimizer_weights").create_dataset("training/SGD/momentum:0", np.zeros(3, np.float32))
    full.attrs["model_config"] = np.bytes_(b'{"class_name": "Model"}')
    back = hdf5.read_keras_weights(full.tobytes())
    assert all(np.array_equal(a, b) for a, b in zip(back, ws)) and len(back) == len(ws)
    with pytest.raises(ValueError):
        hdf5.read_keras_weights(hdf5.Writer().tobytes())
