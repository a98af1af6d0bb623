"""Batch generators with the array contract of the reference's vae/data.py::VAEDataGenerator (lines 23-100):
`__len__`, `__getitem__ -> (M (B,d,d,d,4) float, cond (B,n_bins) one-hot)`, `.list_IDs`, `.batch_size`,
`.list_IDs_temp`, `on_epoch_end`.  Two sources: the reference's on-disk layout
(data/<name>/matrices/{density_matrices,coordinate_grids}/<id>.npy + property csv) and a synthetic source that
voxelises perovskite-like cells on the device (icsg3d_b200.utils.synthetic_batch)."""
from __future__ import annotations

import os
import re

import numpy as np


class VAEDataGenerator:
    def __init__(self, list_IDs, data_path, batch_size=2, dim=(32, 32, 32), n_channels=4, n_classes=95, shuffle=False,
                 property_csv="property.csv", n_bins=10, target="formation_energy_per_atom", return_S=False):
        import pandas as pd
        self.dim, self.batch_size, self.list_IDs = dim, batch_size, list_IDs
        self.n_channels, self.n_classes, self.shuffle, self.data_path = n_channels, n_classes, shuffle, data_path
        self.property_df = pd.read_csv(property_csv)
        self.n_bins = n_bins
        self.property_df["bin"] = pd.qcut(self.property_df[target], self.n_bins, np.arange(n_bins)).astype(int)
        self.return_S = return_S
        self.on_epoch_end()

    def __len__(self):
        return int(np.floor(len(self.list_IDs) / self.batch_size))

    def on_epoch_end(self):
        self.indexes = np.arange(len(self.list_IDs))
        if self.shuffle:
            np.random.shuffle(self.indexes)

    def __getitem__(self, index):
        idx = self.indexes[index * self.batch_size:(index + 1) * self.batch_size]
        self.list_IDs_temp = [self.list_IDs[k] for k in idx]
        M = np.empty((self.batch_size, *self.dim, self.n_channels))
        cond = np.zeros((self.batch_size, self.n_bins))
        if self.return_S:  # vae/data.py:72-86: species grids + binary atom mask ride along with the condition
            S = np.empty((self.batch_size, *self.dim, 1))
            S_b = np.empty((self.batch_size, *self.dim, 1))
        for i, ID in enumerate(self.list_IDs_temp):
            M[i, ..., 0] = np.load(os.path.join(self.data_path, "density_matrices", ID)).reshape(self.dim)
            if self.n_channels > 1:
                M[i, ..., 1:] = np.load(os.path.join(self.data_path, "coordinate_grids", ID)).reshape(*self.dim, 3)
            cif_id = re.split(r"_|\.", ID)[0]
            b = self.property_df[self.property_df["task_id"] == cif_id]["bin"].values
            cond[i, int(b[0])] = 1.0
            if self.return_S:
                S[i] = np.load(os.path.join(self.data_path, "species_matrices", ID)).reshape(*self.dim, 1)
                S_b[i] = np.where(S[i] != 0, 1, 0)
        if self.return_S:
            onehot = np.eye(self.n_classes, dtype=np.float32)[S[..., 0].astype(np.int64)]  # keras.utils.to_categorical
            return M, [cond, onehot, S_b]
        return M, cond


class SyntheticVAEGenerator:
    """Same contract, data generated on the GPU (no files): deterministic per (seed, batch index)."""

    def __init__(self, n_samples, batch_size=20, d=32, n_bins=10, seed=0, device="cuda"):
        self.batch_size, self.d, self.n_bins, self.seed, self.device = batch_size, d, n_bins, seed, device
        self.list_IDs = [f"synthetic-{i}.npy" for i in range(n_samples)]
        self.list_IDs_temp = []

    def __len__(self):
        return len(self.list_IDs) // self.batch_size

    def on_epoch_end(self):
        pass

    def __getitem__(self, index):
        from .. import utils
        self.list_IDs_temp = self.list_IDs[index * self.batch_size:(index + 1) * self.batch_size]
        M, cond, _ = utils.synthetic_batch(self.batch_size, d=self.d, seed=self.seed * 100003 + index, ncond=self.n_bins,
                                           device=self.device)
        return M, cond
