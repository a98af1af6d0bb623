"""Weight files.  The reference stores Keras HDF5 (`save_weights` / `model.save`, lattice_vae.py:339-341,
unet.py:378-379); h5py is not available in this environment, so the native container is a numpy .npz
written AT THE EXACT PATH the reference would use (whatever its extension), holding Keras-layout tensors under
`<layer>/<weight>` names — conv kernels (kd,kh,kw,Cin,Cout), Dense kernels (in,out), BN gamma/beta/
moving_mean/moving_variance.  Importing real Keras .h5 files is a listed next step."""
from __future__ import annotations

import os

import numpy as np


def save_npz(path, tensors: dict):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        np.savez(f, **{k.replace("/", "__"): v for k, v in tensors.items()})


def load_npz(path) -> dict:
    with np.load(path) as z:
        return {k.replace("__", "/"): z[k] for k in z.files}
