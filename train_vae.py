#!/usr/bin/env python
"""Drop-in for the reference's train_vae.py (flags of train_vae.py:30-84) on the B200 path.
Extra flags: --synthetic (on-device voxelised perovskite-like cells instead of data/<name>/matrices), --perceptual
(U-Net weights file; default = the reference's convention, or a seeded random U-Net with --synthetic)."""
import argparse
import os

from icsg3d_b200.vae.data import SyntheticVAEGenerator, VAEDataGenerator
from icsg3d_b200.vae.lattice_vae import LatticeDFCVAE

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--name", type=str, default="synthetic", help="Name of data folder")
    p.add_argument("--samples", type=int, default=40000, help="Total number of training and validation samples")
    p.add_argument("--epochs", type=int, default=50)
    p.add_argument("--batch_size", type=int, default=20)
    p.add_argument("--ncond", type=int, default=10, help="Number of condition bins")
    p.add_argument("--nrot", type=int, default=10, help="Number of augmentations")
    p.add_argument("--cond", type=str, default="formation_energy_per_atom")
    p.add_argument("--split", type=float, default=0.8)
    p.add_argument("--d", type=int, default=32)
    p.add_argument("--synthetic", action="store_true")
    p.add_argument("--perceptual", type=str, default=None)
    a = p.parse_args()
    mode = a.name
    weights_dir = os.path.join("saved_models", "vae", mode)
    os.makedirs(weights_dir, exist_ok=True)
    weights = os.path.join(weights_dir, "vae_weights_" + mode + ".best.hdf5")
    pm = a.perceptual or os.path.join("saved_models", "unet", mode, "unet_weights_" + mode + ".best.h5")
    if a.synthetic:
        n_train = int(a.samples * a.split) // a.batch_size * a.batch_size
        n_val = max(a.batch_size, (a.samples - n_train) // a.batch_size * a.batch_size)
        train_gen = SyntheticVAEGenerator(n_train, a.batch_size, d=a.d, n_bins=a.ncond, seed=1)
        val_gen = SyntheticVAEGenerator(n_val, a.batch_size, d=a.d, n_bins=a.ncond, seed=2)
        if not os.path.exists(pm):
            pm = None
    else:
        from icsg3d_b200.datasplit import data_split
        path = os.path.join("data", mode, "matrices")
        csv_path = os.path.join("data", mode, mode + ".csv")
        tr, va = data_split(path, a.samples, frac=a.split, n_rot=a.nrot)
        tr = tr[: len(tr) // a.batch_size * a.batch_size]
        va = va[: len(va) // a.batch_size * a.batch_size]
        kw = dict(data_path=path, property_csv=csv_path, batch_size=a.batch_size, n_channels=4, shuffle=True, n_bins=a.ncond,
                  target=a.cond)
        train_gen, val_gen = VAEDataGenerator(tr, **kw), VAEDataGenerator(va, **kw)
    vae = LatticeDFCVAE(input_shape=(a.d, a.d, a.d, 4), perceptual_model=pm, cond_shape=a.ncond,
                        output_dir=os.path.join("output", "vae", mode))
    vae.train(train_gen, val_gen, epochs=a.epochs, weights=weights)
