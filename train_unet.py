#!/usr/bin/env python
"""Drop-in for the reference's train_unet.py (flags of train_unet.py:30-78) on the B200 path (+ --synthetic)."""
import argparse
import os

from icsg3d_b200.unet.data import SyntheticUnetGenerator, UnetDataGenerator
from icsg3d_b200.unet.unet import AtomUnet

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--name", type=str, default="synthetic")
    p.add_argument("--samples", type=int, default=20000)
    p.add_argument("--d", type=int, default=32)
    p.add_argument("--epochs", type=int, default=50)
    p.add_argument("--lr", type=float, default=3e-6)
    p.add_argument("--batch_size", type=int, default=10)
    p.add_argument("--nrot", type=int, default=10)
    p.add_argument("--nclasses", type=int, default=95)
    p.add_argument("--split", type=float, default=0.8)
    p.add_argument("--synthetic", action="store_true")
    a = p.parse_args()
    mode = a.name
    wdir = os.path.join("saved_models", "unet", mode)
    os.makedirs(wdir, exist_ok=True)
    weights = os.path.join(wdir, "unet_weights_" + mode + ".best.hdf5")
    if a.synthetic:
        n_train = int(a.samples * a.split) // a.batch_size * a.batch_size
        n_val = max(a.batch_size, (a.samples - n_train) // a.batch_size * a.batch_size)
        train_gen = SyntheticUnetGenerator(n_train, a.batch_size, d=a.d, seed=1)
        val_gen = SyntheticUnetGenerator(n_val, a.batch_size, d=a.d, seed=2)
    else:
        from icsg3d_b200.datasplit import data_split
        path = os.path.join("data", mode, "matrices")
        tr, va = data_split(path, a.samples, frac=a.split, n_rot=a.nrot)
        train_gen = UnetDataGenerator(tr, data_path=path, batch_size=a.batch_size, n_channels=4, shuffle=True, compact_labels=True)
        val_gen = UnetDataGenerator(va, data_path=path, batch_size=a.batch_size, n_channels=4, shuffle=True, compact_labels=True)
    unet = AtomUnet(num_classes=a.nclasses, weights=weights, input_shape=(a.d, a.d, a.d, 4), lr=a.lr)
    unet.train_generator(train_gen, val_gen, epochs=a.epochs, output_dir=os.path.join("output", "unet", mode))
    unet.save_(weights, os.path.join(wdir, "unet_weights_" + mode + ".best.h5"))
