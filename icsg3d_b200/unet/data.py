"""Batch generators with the contract of the reference's unet/data.py::UnetDataGenerator (lines 20-100):
`__getitem__ -> (X (B,d,d,d,C), [one-hot species (B,d,d,d,95), mask (B,d,d,d,1)])`.  The B200 path also accepts the
compact form `(X, uint8 species)` (the one-hot tensor is 12.4 MB per sample and is never materialised on the device)."""
from __future__ import annotations

import os

import numpy as np


class UnetDataGenerator:
    def __init__(self, list_IDs, data_path, batch_size=2, dim=(32, 32, 32), n_channels=7, n_classes=95, shuffle=False,
                 compact_labels=False):
        self.dim, self.batch_size, self.list_IDs = dim, batch_size, list_IDs
        self.n_channels, self.n_classes, self.shuffle, self.data_path = n_channels, n_classes, shuffle, data_path
        self.compact_labels = compact_labels
        self.on_epoch_end()

    def __len__(self):
        return int(np.floor(len(self.list_IDs) / self.batch_size))

    def on_epoch_end(self):
        self.indexes = np.arange(len(self.list_IDs))
        if self.shuffle:
            np.random.shuffle(self.indexes)

    def __getitem__(self, index):
        idx = self.indexes[index * self.batch_size:(index + 1) * self.batch_size]
        ids = [self.list_IDs[k] for k in idx]
        X = np.empty((self.batch_size, *self.dim, self.n_channels))
        S = np.empty((self.batch_size, *self.dim), dtype=np.uint8)
        for i, ID in enumerate(ids):
            X[i, ..., 0] = np.load(os.path.join(self.data_path, "density_matrices", ID)).reshape(self.dim)
            if self.n_channels > 1:
                X[i, ..., 1:] = np.load(os.path.join(self.data_path, "coordinate_grids", ID)).reshape(*self.dim, 3)
            S[i] = np.load(os.path.join(self.data_path, "species_matrices", ID)).reshape(self.dim).astype(np.uint8)
        if self.compact_labels:
            return X, S
        onehot = np.eye(self.n_classes, dtype=np.float32)[S]
        return X, [onehot, (S != 0).astype(np.float32)[..., None]]


class SyntheticUnetGenerator:
    def __init__(self, n_samples, batch_size=10, d=32, seed=0, device="cuda"):
        self.batch_size, self.d, self.seed, self.device = batch_size, d, seed, device
        self.list_IDs = [f"synthetic-{i}.npy" for i in range(n_samples)]

    def __len__(self):
        return len(self.list_IDs) // self.batch_size

    def on_epoch_end(self):
        pass

    def __getitem__(self, index):
        from .. import utils
        M, _, S = utils.synthetic_batch(self.batch_size, d=self.d, seed=self.seed * 100003 + index, device=self.device)
        return M, S
