"""Golden vectors for the post-processing / augmentation rows from the REFERENCE ITSELF
(/root/reference/utils.py: to_lattice_params, to_voxel_params, random_rotation_3d).

Run in the build container only:  python tests/golden/make_post_golden.py
Inputs are taken from tests/golden/voxel_golden.npz (the reference voxeliser's own outputs: density M, species S,
coordinate grid p per case), so only the reference OUTPUTS are stored here:

  lattice:  for float64 and float32 stacks of the d=32 coordinate grids (plus a seeded ReLU'd perturbation that mimics
            a decoder output) -> lp, dv as the reference returns them (dtype preserved);
  rotation: for every case k, np.random.seed(100+k); random_rotation_3d(M, S, p) -> the drawn axis sequence, S_rot
            (uint8, exact) for all cases and M_rot / p_rot (float64) for the d=16 cases.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference_utils  # noqa: E402


def lattice_inputs(z):
    """The stacks both this script and tests/test_gpu_post.py build from voxel_golden.npz."""
    ks = [k for k in range(20) if f"c{k}_p" in z.files and z[f"c{k}_p"].shape[0] == 32]
    p64 = np.stack([z[f"c{k}_p"] for k in ks])                       # (B,32,32,32,3) float64, as utils.coordinate_grid returns
    rng = np.random.default_rng(5)
    noisy = np.maximum(p64 + rng.normal(0, 0.05, size=p64.shape), 0.0)   # decoder-like: ReLU output, not an exact grid
    return {"grid64": p64, "grid32": p64.astype(np.float32), "noisy32": noisy.astype(np.float32), "noisy64": noisy}


def main():
    u = load_reference_utils()
    z = np.load(os.path.join(ROOT, "tests", "golden", "voxel_golden.npz"))
    out = {}
    for name, p in lattice_inputs(z).items():
        lp = u.to_lattice_params(p)
        dv = u.to_voxel_params(lp)
        out[f"lat_{name}_lp"], out[f"lat_{name}_dv"] = lp, dv
        print(name, lp.dtype, lp[0], dv[0])
    k = 0
    while f"c{k}_M" in z.files:
        M, S, p = z[f"c{k}_M"], z[f"c{k}_S"].astype(np.float64), z[f"c{k}_p"]
        np.random.seed(100 + k)
        state = np.random.get_state()
        seq = np.random.choice(3, 3)
        np.random.set_state(state)
        Mr, Sr, pr = u.random_rotation_3d(M, S, p)
        out[f"rot{k}_seq"] = seq.astype(np.int64)
        out[f"rot{k}_S"] = Sr.astype(np.uint8)
        assert np.array_equal(Sr, Sr.astype(np.uint8))
        if M.shape[0] == 16:
            out[f"rot{k}_M"], out[f"rot{k}_p"] = Mr, pr
        print("rot", k, M.shape, seq)
        k += 1
    path = os.path.join(ROOT, "tests", "golden", "post_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
