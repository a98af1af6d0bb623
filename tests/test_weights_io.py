"""Keras HDF5 weight import/export without h5py (icsg3d_b200/h5lite.py, weights_io.py).

The reader is pinned on a file written by libhdf5 itself (tests/golden/libhdf5_written_sample.h5 = scipy's MATLAB v7.3
test file: 512-byte user block, superblock v0, v1 object headers, symbol-table group, layout message v2, attribute);
the writer is checked by round trips through that reader, laid out like Keras 2.3.1 `save_weights` files
(lattice_vae.py:339-341, unet.py:378-379)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "libhdf5_written_sample.h5")


def test_reader_on_libhdf5_written_file():
    from icsg3d_b200 import h5lite
    assert h5lite.is_hdf5(GOLD)
    root = h5lite.read(GOLD)
    assert root.keys() == ["testdouble"]
    d = root["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.float64
    assert np.array_equal(d.ravel(), np.arange(9) * np.pi / 4)


def test_h5_roundtrip_many_links_and_attrs(tmp_path):
    from icsg3d_b200 import h5lite
    rng = np.random.default_rng(0)
    root = h5lite.Node(attrs={"backend": b"tensorflow", "n": np.int64(7), "v": np.arange(3, dtype=np.float64)})
    ws = {}
    for i in range(70):  # > 8 links: several SNOD nodes under one B-tree node
        ws[i] = rng.standard_normal((2, 3, i % 4 + 1)).astype(np.float32)
        root.group("layers").group("layer_%02d" % i).items["kernel:0"] = ws[i]
    root.group("layers").attrs["weight_names"] = np.array([b"a/kernel:0", b"bb/bias:0"])
    root.items["empty"] = np.zeros((0, 4), np.float32)
    p = str(tmp_path / "t.h5")
    h5lite.write(p, root)
    back = h5lite.read(p)
    assert back.attrs["backend"] == b"tensorflow" and back.attrs["n"] == 7 and np.array_equal(back.attrs["v"], [0, 1, 2])
    assert sorted(back["layers"].keys()) == ["layer_%02d" % i for i in range(70)]
    assert list(back["layers"].attrs["weight_names"]) == [b"a/kernel:0", b"bb/bias:0"]
    for i in range(70):
        assert np.array_equal(back["layers/layer_%02d/kernel:0" % i], ws[i])
    assert back["empty"].shape == (0, 4)


def _rand_tensors(specs, seed):
    rng = np.random.default_rng(seed)
    return {n: rng.standard_normal(sh).astype(np.float32) for n, sh, _, _ in specs}


@pytest.mark.parametrize("model", ["unet", "vae"])
def test_keras_h5_export_import(tmp_path, model):
    from icsg3d_b200 import weights_io
    from icsg3d_b200.params import unet_specs, vae_specs
    specs = unet_specs(4, 95) if model == "unet" else vae_specs()
    t = _rand_tensors(specs, 1)
    p = str(tmp_path / "sub" / f"{model}_weights_x.best.hdf5")
    weights_io.save_weights_file(p, t, specs, model=model)
    from icsg3d_b200 import h5lite
    assert h5lite.is_hdf5(p)
    back = weights_io.load_weights_file(p, specs)
    assert set(back) == set(t) and all(np.array_equal(back[k], t[k]) for k in t)
    root = h5lite.read(p)
    if model == "vae":  # nested sub-models, trainable weights first (Keras Network.weights)
        assert list(root.attrs["layer_names"]) == [b"encoder", b"decoder"]
        names = [n.decode() for n in root["encoder"].attrs["weight_names"]]
        first_state = next(i for i, n in enumerate(names) if "moving" in n)
        assert all("moving" in n for n in names[first_state:]) and names[0] == "enc_conv1/kernel:0"
    else:
        assert root.attrs["layer_names"][0] == b"c1" and "c1" in root["c1"].keys()


def test_keras_h5_import_with_foreign_layer_names(tmp_path):
    """A file as Keras itself names things (conv3d_7, batch_normalization_3, full-model `model_weights/` group, heads in
    the other order) must map onto our specs by order + shape."""
    from icsg3d_b200 import h5lite, weights_io
    from icsg3d_b200.params import unet_specs
    specs = unet_specs(4, 95)
    t = _rand_tensors(specs, 2)
    layers = []
    for n, _, _, _ in specs:
        l = n.split("/")[0]
        if l not in layers:
            layers.append(l)
    layers[-2], layers[-1] = layers[-1], layers[-2]  # sig before soft
    root = h5lite.Node(attrs={"keras_version": b"2.3.1"})
    mw = root.group("model_weights")
    knames = []
    for i, l in enumerate(layers):
        kn = (f"batch_normalization_{i}" if l.startswith("bn_") else f"conv3d_{i + 5}")
        knames.append(kn)
        g = mw.group(kn)
        wn = []
        for n, _, _, _ in specs:
            if n.split("/")[0] == l:
                w = n.split("/")[1]
                g.group(kn).items[w + ":0"] = t[n]
                wn.append(f"{kn}/{w}:0".encode())
        g.attrs["weight_names"] = np.array(wn)
    mw.group("input_1").attrs["weight_names"] = np.zeros((0,), "S1")
    mw.attrs["layer_names"] = np.array([b"input_1"] + [k.encode() for k in knames])
    p = str(tmp_path / "unet.h5")
    h5lite.write(p, root)
    back = weights_io.load_weights_file(p, specs)
    assert all(np.array_equal(back[k], t[k]) for k in t)
    # a file for a different architecture fails loudly
    with pytest.raises((ValueError, KeyError)):
        weights_io.load_weights_file(p, unet_specs(1, 95))


def test_npz_container_and_magic_detection(tmp_path):
    from icsg3d_b200 import weights_io
    d = {"enc_conv1/kernel": np.random.rand(3, 3, 3, 4, 2).astype(np.float32), "enc_bn1/gamma": np.ones(2, np.float32)}
    p = str(tmp_path / "w.npz")
    weights_io.save_weights_file(p, d)
    back = weights_io.load_weights_file(p)
    assert set(back) == set(d) and all(np.array_equal(back[k], d[k]) for k in d)
    junk = tmp_path / "junk.hdf5"
    junk.write_bytes(b"not a weight file")
    with pytest.raises(ValueError):
        weights_io.load_weights_file(str(junk))
