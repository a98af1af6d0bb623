"""Weight files of the drop-in classes.

The reference stores Keras HDF5 (`save_weights` / `model.save`: lattice_vae.py:149-151,339-341, unet.py:261-264,
378-379).  Two containers are understood, told apart by their magic bytes (never by the file extension):

* **Keras HDF5** (`\\x89HDF`) — read and written with the pure-Python `h5lite` (no h5py in this image).  Import accepts
  weights-only files and full-model files (`model_weights/` group), flat models (one group per layer: the U-Net) and
  nested ones (the VAE's `encoder` / `decoder` sub-models, whose weights Keras lists trainable-first).  Tensors are
  matched to our parameter specs the way Keras' own `load_weights` does — by ORDER within each weight kind (kernel, bias,
  gamma, beta, moving_mean, moving_variance) — with a shape check, because Keras' auto-generated layer names
  (`conv3d_7`, `batch_normalization_3`, ...) depend on what else was built in the writing process.
  Export writes the same structure under our layer names; the reference's `model.load_weights(path)` (order-based)
  accepts it.  Full-model files (`load_model`) need Keras' JSON model config and are not written.
* the native **.npz** (`PK`): Keras-layout tensors under `<layer>/<weight>` names.

`save_weights_file` picks HDF5 for `.h5` / `.hdf5` paths (the paths the reference uses) and .npz otherwise.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

from . import h5lite

KINDS = ("kernel", "bias", "gamma", "beta", "moving_mean", "moving_variance")


def save_npz(path, tensors: dict):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        np.savez(f, **{k.replace("/", "__"): v for k, v in tensors.items()})


def load_npz(path) -> dict:
    with np.load(path) as z:
        return {k.replace("__", "/"): z[k] for k in z.files}


# ------------------------------------------------------------------------------------------------------
# Keras HDF5
# ------------------------------------------------------------------------------------------------------
def _attr_names(group, key):
    """Keras splits long name lists into key0, key1, ... (saving.py HDF5_OBJECT_HEADER_LIMIT)."""
    if key in group.attrs:
        vals = [group.attrs[key]]
    else:
        vals, i = [], 0
        while f"{key}{i}" in group.attrs:
            vals.append(group.attrs[f"{key}{i}"])
            i += 1
    out = []
    for v in vals:
        for n in np.atleast_1d(v):
            out.append(n.decode("utf8") if isinstance(n, bytes) else str(n))
    return out


def keras_h5_tensors(path):
    """-> [(layer, weight_name, kind, array)] in file order (layer_names x weight_names)."""
    root = h5lite.read(path)
    g = root["model_weights"] if "model_weights" in root else root
    out = []
    for layer in _attr_names(g, "layer_names"):
        lg = g[layer]
        for wn in _attr_names(lg, "weight_names"):
            kind = wn.split("/")[-1].split(":")[0]
            out.append((layer, wn, kind, np.asarray(lg[wn])))
    return out


def match_keras_tensors(tensors, specs, strict=True) -> dict:
    """Assign file tensors to spec names: k-th tensor of a kind -> first unassigned spec of that kind with the same
    shape (Keras' order-based loading, robust to auto-generated layer names and to the trainable-first order of nested
    models).  strict: every spec must be filled and every file tensor used."""
    by_kind = {k: [] for k in KINDS}
    for name, shape, _, _ in specs:
        kind = name.split("/")[-1]
        if kind in by_kind:
            by_kind[kind].append((name, tuple(shape)))
    used, out = set(), OrderedDict()
    for layer, wn, kind, arr in tensors:
        cands = [(n, sh) for n, sh in by_kind.get(kind, []) if n not in used]
        hit = next((n for n, sh in cands if sh == tuple(arr.shape)), None)
        if hit is None:
            if strict:
                raise ValueError(f"Keras weight {layer}/{wn} {tuple(arr.shape)} has no matching parameter left "
                                 f"(next expected {kind}: {cands[0] if cands else 'none'})")
            continue
        used.add(hit)
        out[hit] = arr.astype(np.float32)
    if strict:
        missing = [n for k in KINDS for n, _ in by_kind[k] if n not in used]
        if missing:
            raise KeyError(f"Keras weight file lacks {len(missing)} tensors, first: {missing[:4]}")
    return out


def _keras_tree(tensors: dict, specs, model):
    """h5lite.Node tree in Keras' save_weights layout.  model='unet': one group per layer; 'vae': the nested `encoder` /
    `decoder` sub-models with their weights listed trainable-first (keras/engine/saving.py)."""
    root = h5lite.Node(attrs={"backend": b"tensorflow", "keras_version": b"2.3.1"})
    layers = OrderedDict()
    for name, _, trainable, _ in specs:
        layer, w = name.split("/")
        layers.setdefault(layer, []).append((w, trainable))
    if model == "vae":
        split = list(layers).index("dec_dense")
        groups = OrderedDict(encoder=list(layers)[:split], decoder=list(layers)[split:])
        for gname, members in groups.items():
            g = root.group(gname)
            names = []
            for want_trainable in (True, False):
                for layer in members:
                    for w, tr in layers[layer]:
                        if tr == want_trainable:
                            names.append(f"{layer}/{w}:0")
                            g.group(layer).items[w + ":0"] = np.asarray(tensors[f"{layer}/{w}"], dtype=np.float32)
            g.attrs["weight_names"] = np.array([n.encode() for n in names])
        root.attrs["layer_names"] = np.array([b"encoder", b"decoder"])
    else:
        for layer, ws in layers.items():
            g = root.group(layer)
            for w, _ in ws:
                g.group(layer).items[w + ":0"] = np.asarray(tensors[f"{layer}/{w}"], dtype=np.float32)
            g.attrs["weight_names"] = np.array([f"{layer}/{w}:0".encode() for w, _ in ws])
        root.attrs["layer_names"] = np.array([l.encode() for l in layers])
    return root


def save_keras_h5(path, tensors: dict, specs, model="unet"):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    h5lite.write(path, _keras_tree(tensors, specs, model))


# ------------------------------------------------------------------------------------------------------
# front door
# ------------------------------------------------------------------------------------------------------
def load_weights_file(path, specs=None) -> dict:
    """{spec name: array} from a Keras HDF5 file (needs `specs`) or the native .npz."""
    if h5lite.is_hdf5(path):
        if specs is None:
            raise ValueError("loading a Keras HDF5 file needs the parameter specs to map its tensors")
        return match_keras_tensors(keras_h5_tensors(path), specs)
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic != b"PK":
        raise ValueError(f"{path}: neither a Keras HDF5 file nor an icsg3d .npz weight container")
    return load_npz(path)


def save_weights_file(path, tensors: dict, specs=None, model="unet"):
    if specs is not None and os.path.splitext(path)[1].lower() in (".h5", ".hdf5"):
        save_keras_h5(path, tensors, specs, model)
    else:
        save_npz(path, tensors)
