"""Generate golden vectors for the voxeliser from the REFERENCE ITSELF (/root/reference/utils.py).

Run in the build container only:  python tests/golden/make_voxel_golden.py
Writes tests/golden/voxel_golden.npz: for each seeded cell the inputs (N, z, l, sigma, d, label_frac,
eps_frac) and the reference outputs M (fp64), S (stored as uint8 — integer valued) and coordinate_grid p.
Cases: perovskite-like ABX3 cells (SURVEY §8c LaFeO3-like probe first), random multi-site cells with many
contested voxels, a single atom, a cell whose spheres do not reach any voxel, d=16 and d=32.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference_utils  # noqa: E402
from oracle import voxelizer as vox  # noqa: E402


def cases():
    rng = np.random.default_rng(0)
    out = []
    a = 3.93
    frac = vox.ABX3_FRAC
    out.append(dict(N=frac * a, z=np.array([57, 26, 8, 8, 8.0]), l=np.array([a, a, a]),
                    sigma=np.array([1.17, 0.72, 1.26, 1.26, 1.26]), d=32, label_frac=1.0, eps_frac=0.25))
    for _ in range(3):
        N, z, l, s = vox.synthetic_cell(rng)
        out.append(dict(N=N, z=z, l=l, sigma=s, d=32, label_frac=1.0, eps_frac=0.25))
    for n in (1, 12, 40):
        l = rng.uniform(3.0, 9.0, size=3)
        N = rng.uniform(0, 1, size=(n, 3)) * l
        z = rng.integers(1, 95, size=n).astype(np.float64)
        s = rng.uniform(0.5, 1.6, size=n)
        out.append(dict(N=N, z=z, l=l, sigma=s, d=16, label_frac=1.0, eps_frac=0.25))
    N, z, l, s = vox.synthetic_cell(rng)
    out.append(dict(N=N, z=z, l=l, sigma=s * 0.7, d=16, label_frac=0.6, eps_frac=0.1))
    out.append(dict(N=np.array([[1.0, 1.0, 1.0]]), z=np.array([8.0]), l=np.array([4.0, 5.0, 6.0]),
                    sigma=np.array([0.01]), d=16, label_frac=1.0, eps_frac=0.25))
    return out


def main():
    ref = load_reference_utils()
    blob = {}
    cs = cases()
    for i, c in enumerate(cs):
        dims = (c["d"],) * 3
        M, S = ref.density_matrix(c["N"], c["z"].copy(), c["l"], dims=dims, sigma=c["sigma"],
                                  label_frac=c["label_frac"], eps_frac=c["eps_frac"])
        p = ref.coordinate_grid(c["l"], dim=c["d"], eps_frac=c["eps_frac"])
        assert np.array_equal(S, np.rint(S)) and S.min() >= 0 and S.max() < 256
        for k in ("N", "z", "l", "sigma"):
            blob[f"c{i}_{k}"] = np.asarray(c[k], dtype=np.float64)
        blob[f"c{i}_meta"] = np.array([c["d"], c["label_frac"], c["eps_frac"]], dtype=np.float64)
        blob[f"c{i}_M"] = M
        blob[f"c{i}_S"] = S.astype(np.uint8)
        blob[f"c{i}_p"] = p.astype(np.float64)
        print(i, c["d"], len(c["z"]), "S hist", dict(zip(*np.unique(S, return_counts=True))) if len(c["z"]) <= 5 else "...",
              "M range", M.min(), M.max())
    blob["ncases"] = np.array([len(cs)])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "voxel_golden.npz"), **blob)


if __name__ == "__main__":
    main()
