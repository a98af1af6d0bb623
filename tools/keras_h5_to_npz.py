#!/usr/bin/env python
"""Offline converter: any Keras .h5/.hdf5 weight file (also chunked / compressed / new-style HDF5 that the in-tree
h5lite reader does not cover) -> the native .npz container of icsg3d_b200.weights_io.  Run it where h5py is installed:

    python tools/keras_h5_to_npz.py saved_models/unet/perov/unet_weights_perov.best.h5 unet --channels 4 -o unet.npz
    python tools/keras_h5_to_npz.py saved_models/vae/perov/vae_weights_perov.best.hdf5 vae -o vae.npz

Tensor matching is the one weights_io.match_keras_tensors implements (order within each weight kind + shape check),
i.e. what Keras' own order-based `load_weights` does (reference: lattice_vae.py:149-151, unet.py:261-264)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def tensors_with_h5py(path):
    import h5py
    out = []
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f
        for layer in [n.decode() if isinstance(n, bytes) else n for n in g.attrs["layer_names"]]:
            lg = g[layer]
            for wn in [n.decode() if isinstance(n, bytes) else n for n in lg.attrs["weight_names"]]:
                out.append((layer, wn, wn.split("/")[-1].split(":")[0], np.asarray(lg[wn])))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("model", choices=["unet", "vae"])
    ap.add_argument("--channels", type=int, default=4)
    ap.add_argument("--classes", type=int, default=95)
    ap.add_argument("--d", type=int, default=32)
    ap.add_argument("-o", "--out", required=True)
    a = ap.parse_args()
    from icsg3d_b200 import weights_io
    from icsg3d_b200.params import unet_specs, vae_specs
    specs = unet_specs(a.channels, a.classes) if a.model == "unet" else vae_specs(a.channels, 10, a.d)
    try:
        tensors = tensors_with_h5py(a.src)
    except ImportError:
        tensors = weights_io.keras_h5_tensors(a.src)  # classic-format files only
    weights_io.save_npz(a.out, weights_io.match_keras_tensors(tensors, specs))
    print("wrote", a.out)
