#!/usr/bin/env python
"""The GPU part of the reference's generate.py (lines 172-225): encode a base compound, sample around it, decode,
lattice parameters from the coordinate channels, U-Net segmentation, argmax / 0.8-threshold.  The per-sample CPU tail
of the reference (watershed, CIF writing, CGCNN; generate.py:228-314) is out of scope (SURVEY §2 #9,#10).
With --synthetic the base compound is a voxelised perovskite-like cell and the networks are seeded random."""
import argparse
import os
import time

import numpy as np

from icsg3d_b200 import utils
from icsg3d_b200.pipeline import GeneratePipeline
from icsg3d_b200.unet.unet import AtomUnet
from icsg3d_b200.vae.lattice_vae import LatticeDFCVAE

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--name", type=str, default="synthetic")
    p.add_argument("--base", type=str, default="LaFeO3")
    p.add_argument("--batch_size", type=int, default=100)
    p.add_argument("--nsamples", type=int, default=100)
    p.add_argument("--var", type=float, default=0.5)
    p.add_argument("--eps_frac", type=float, default=0.25)
    p.add_argument("--ncond", type=int, default=10)
    p.add_argument("--d", type=int, default=32)
    p.add_argument("--synthetic", action="store_true")
    p.add_argument("--out", type=str, default=None)
    a = p.parse_args()
    mode, d = a.name, a.d
    vae_w = os.path.join("saved_models", "vae", mode, "vae_weights_" + mode + ".best.hdf5")
    unet_w = os.path.join("saved_models", "unet", mode, "unet_weights_" + mode + ".best.hdf5")
    pm = os.path.join("saved_models", "unet", mode, "unet_weights_" + mode + ".best.h5")
    vae = LatticeDFCVAE(input_shape=(d, d, d, 4), perceptual_model=pm if os.path.exists(pm) else None, cond_shape=a.ncond)
    vae._set_model(vae_w, batch_size=a.batch_size)
    unet = AtomUnet(weights=unet_w, input_shape=(d, d, d, 4))
    if a.synthetic:
        M_base, cond, _ = utils.synthetic_batch(1, d=d, seed=7, ncond=a.ncond)
        M_base, cond = M_base.cpu().numpy(), cond.cpu().numpy()
    else:
        import pandas as pd
        path = os.path.join("data", mode, "matrices")
        df = pd.read_csv(os.path.join("data", mode, mode + ".csv"))
        df["interval"] = pd.qcut(df["formation_energy_per_atom"], a.ncond, np.arange(a.ncond))
        base = df[df["pretty_formula"] == a.base]["task_id"].values[0] if not a.base.startswith("mp-") else a.base
        Mb = np.load(os.path.join(path, "density_matrices", base + ".npy")).reshape(1, d, d, d, 1)
        Cb = np.load(os.path.join(path, "coordinate_grids", base + ".npy")).reshape(1, d, d, d, 3)
        M_base = np.concatenate([Mb, Cb], axis=-1)
        cond = np.eye(a.ncond, dtype=np.float32)[[int(df[df["task_id"] == base]["interval"].values[0])]]
    z_mu, z_logvar, z = vae.encoder.predict([M_base, cond])
    # decode -> lattice parameters -> U-Net -> argmax / 0.8 threshold, device resident (icsg3d_b200/pipeline.py): the decoder
    # output never visits the host; only labels, mask, density channel and lattice / voxel parameters come back
    pipe = GeneratePipeline(vae, unet, a.batch_size, threshold=0.8, eps_frac=a.eps_frac)
    cond_tensor = np.tile(cond, (a.batch_size, 1))
    t0 = time.time()
    n_done = 0
    for b in range(a.nsamples // a.batch_size):
        z_s = np.random.normal(z_mu, a.var, size=(a.batch_size, vae.latent_dim))   # generate.py:204 (`var` used as a std)
        r = pipe.run(z_s, cond_tensor)                                              # generate.py:208-225
        S_prime, S_b = r["species"].cpu().numpy(), r["mask"].cpu().numpy().astype(bool)
        l_prime, dv = r["lattice"].cpu().numpy(), r["voxel"].cpu().numpy()
        n_done += a.batch_size
        if a.out:
            os.makedirs(a.out, exist_ok=True)
            np.savez_compressed(os.path.join(a.out, f"batch_{b}.npz"), M=r["density"].cpu().numpy(), S=S_prime, mask=S_b,
                                lattice=l_prime, dv=dv)
    dt = time.time() - t0
    print("generated %d samples in %.2f s (%.1f samples/s): decoded densities, species labels, atom masks, lattice params"
          % (n_done, dt, n_done / max(dt, 1e-9)))
