"""Pins the voxeliser oracle (oracle/voxelizer.py) to golden vectors produced by the REFERENCE itself
(/root/reference/utils.py::density_matrix, coordinate_grid — tests/golden/make_voxel_golden.py)."""
import os

import numpy as np
import pytest

from oracle import voxelizer as vox

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxel_golden.npz"))
NCASES = int(GOLD["ncases"][0])


def case(i):
    d, label_frac, eps_frac = GOLD[f"c{i}_meta"]
    return dict(N=GOLD[f"c{i}_N"], z=GOLD[f"c{i}_z"], l=GOLD[f"c{i}_l"], sigma=GOLD[f"c{i}_sigma"], d=int(d),
                label_frac=float(label_frac), eps_frac=float(eps_frac), M=GOLD[f"c{i}_M"], S=GOLD[f"c{i}_S"],
                p=GOLD[f"c{i}_p"])


@pytest.mark.parametrize("i", range(NCASES))
def test_species_grid_bit_exact(i):
    c = case(i)
    M, S = vox.density_matrix(c["N"], c["z"], c["l"], dims=(c["d"],) * 3, sigma=c["sigma"], label_frac=c["label_frac"],
                              eps_frac=c["eps_frac"])
    assert np.array_equal(S, c["S"].astype(np.float64))
    np.testing.assert_allclose(M, c["M"], rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("i", range(NCASES))
def test_coordinate_grid_bit_exact(i):
    c = case(i)
    p = vox.coordinate_grid(c["l"], dim=c["d"], eps_frac=c["eps_frac"])
    assert np.array_equal(p, c["p"])


def test_survey_histogram_pin():
    """LaFeO3-like cell, a=3.93 -> S histogram {0:27570, 8:3900, 26:280, 57:1018}."""
    c = case(0)
    vals, counts = np.unique(c["S"], return_counts=True)
    assert dict(zip(vals.tolist(), counts.tolist())) == {0: 27570, 8: 3900, 26: 280, 57: 1018}
    assert abs(c["M"].max() - 4.615) < 1e-3 and abs(c["M"].mean() - 0.3574) < 1e-4


def test_lattice_params_quirk():
    """to_lattice_params returns a*(1-1/d): a=4 -> 3.875 at d=32."""
    p = vox.coordinate_grid([4.0, 4.0, 4.0], dim=32)[None]
    lp = vox.to_lattice_params(p)
    np.testing.assert_allclose(lp, [[3.875, 3.875, 3.875]], rtol=1e-12)
