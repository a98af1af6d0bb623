"""T4: the CUDA voxeliser through the drop-in utils API against the golden vectors produced by the reference
itself (tests/golden/voxel_golden.npz) — species grid S bit-exact, density M to 1e-12, coordinate grid to fp32."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxel_golden.npz"))
NCASES = int(GOLD["ncases"][0])


@pytest.mark.parametrize("i", range(NCASES))
def test_density_matrix_matches_reference(i):
    from icsg3d_b200 import utils
    d, label_frac, eps_frac = GOLD[f"c{i}_meta"]
    d = int(d)
    M, S = utils.density_matrix(GOLD[f"c{i}_N"], GOLD[f"c{i}_z"], GOLD[f"c{i}_l"], dims=(d, d, d), sigma=GOLD[f"c{i}_sigma"],
                                label_frac=float(label_frac), eps_frac=float(eps_frac))
    assert S.dtype == np.float64 and M.dtype == np.float64
    assert np.array_equal(S, GOLD[f"c{i}_S"].astype(np.float64)), "species grid must be bit-exact"
    np.testing.assert_allclose(M, GOLD[f"c{i}_M"], rtol=1e-12, atol=1e-300)
    p = utils.coordinate_grid(GOLD[f"c{i}_l"], dim=d, eps_frac=float(eps_frac))
    np.testing.assert_allclose(p, GOLD[f"c{i}_p"], rtol=1e-6, atol=1e-7)


def test_synthetic_batch_equals_oracle_on_same_cells():
    """Device generator -> voxeliser, re-checked cell by cell with the numpy oracle."""
    import torch
    from icsg3d_b200 import utils
    from oracle import voxelizer as vox
    sites, nsites, lat = utils.synthetic_cells(6, seed=5)
    m32, m64, s8, _ = utils.voxelize_cells(sites, nsites, lat, d=32, want_m64=True)
    torch.cuda.synchronize()
    sites, lat = sites.cpu().numpy(), lat.cpu().numpy()
    for c in range(6):
        n = int(nsites[c])
        N, z, sigma = sites[c, :n, 0:3], sites[c, :n, 6], sites[c, :n, 7]
        M, S = vox.density_matrix(N, z, lat[c], dims=(32, 32, 32), sigma=sigma)
        assert np.array_equal(s8[c].cpu().numpy(), S.astype(np.uint8))
        np.testing.assert_allclose(m64[c].cpu().numpy(), M, rtol=1e-11)
        np.testing.assert_allclose(m32[c, ..., 0].cpu().numpy(), M.astype(np.float32), rtol=1e-6)
        np.testing.assert_allclose(m32[c, ..., 1:].cpu().numpy(), vox.coordinate_grid(lat[c]).astype(np.float32), rtol=1e-6)
    assert 3.7 <= lat.min() and lat.max() <= 4.3


@pytest.mark.parametrize("i", range(NCASES))
def test_fast_network_path_species_bit_exact_density_fp32(i):
    """The fast kernel behind the network outputs (fp32 input tensor + uint8 species, no fp64 outputs requested):
    squared-distance decisions with an exact-sqrt guard band -> species still BIT-EXACT against the reference's golden
    grids; fp64 range reduction + SFU 2^r for the Gaussian -> density within a few fp32 ulp of the rounded fp64 value."""
    import torch
    from icsg3d_b200 import utils
    d, label_frac, eps_frac = GOLD[f"c{i}_meta"]
    d = int(d)
    rec = utils.site_records(GOLD[f"c{i}_N"], GOLD[f"c{i}_z"], GOLD[f"c{i}_sigma"], float(label_frac))[None]
    sites = torch.from_numpy(rec).cuda()
    nsites = torch.tensor([rec.shape[1]], dtype=torch.int32, device="cuda")
    lat = torch.tensor(np.asarray(GOLD[f"c{i}_l"], dtype=np.float64)[:3].reshape(1, 3), device="cuda")
    m32, m64, s8, s64 = utils.voxelize_cells(sites, nsites, lat, d=d, eps_frac=float(eps_frac))
    assert m64 is None and s64 is None
    assert np.array_equal(s8[0].cpu().numpy(), GOLD[f"c{i}_S"]), "species grid must be bit-exact on the fast path too"
    want = GOLD[f"c{i}_M"].astype(np.float32)
    np.testing.assert_allclose(m32[0, ..., 0].cpu().numpy(), want, rtol=1e-6, atol=1e-37)
    np.testing.assert_allclose(m32[0, ..., 1:].cpu().numpy(), GOLD[f"c{i}_p"].astype(np.float32), rtol=1e-6, atol=1e-7)
