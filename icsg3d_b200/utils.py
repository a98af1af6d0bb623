"""Drop-in for the voxeliser functions of the reference's utils.py, running on the GPU.

    density_matrix(N, z, l, dims, sigma, dist, label_frac, eps_frac) -> (M, S)     utils.py:97-144
    coordinate_grid(l, dim, eps_frac)                                              utils.py:88-94
    to_lattice_params / to_voxel_params                                            utils.py:160-190

plus the batched device generator used for synthetic benchmark inputs.  The heavy part (d^3 x nsites
distances, Gaussians, species rule) runs in libicsg3d.so::icsg3d_voxelize; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

MAX_SITES = 64
_REC = 8


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def site_records(N, z, sigma, label_frac=1.0):
    """Per-site constants exactly as numpy computes them in utils.py:122,136-137."""
    N = np.asarray(N, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64) * np.ones_like(z)
    rec = np.zeros((len(z), _REC), dtype=np.float64)
    rec[:, 0:3] = N
    rec[:, 3] = sigma * label_frac
    rec[:, 4] = z / (sigma ** 3)
    rec[:, 5] = 2 * sigma ** 2
    rec[:, 6] = z
    rec[:, 7] = sigma
    return rec


def voxelize_cells(sites, nsites, lattice, d=32, eps_frac=0.25, want_m32=True, want_m64=False, want_s8=True,
                   want_s64=False):
    """sites: cuda fp64 [ncells,max_sites,8]; nsites: cuda int32 [ncells]; lattice: cuda fp64 [ncells,3]."""
    ncells, max_sites = sites.shape[0], sites.shape[1]
    dev = sites.device
    m32 = torch.empty(ncells, d, d, d, 4, dtype=torch.float32, device=dev) if want_m32 else None
    m64 = torch.empty(ncells, d, d, d, dtype=torch.float64, device=dev) if want_m64 else None
    s8 = torch.empty(ncells, d, d, d, dtype=torch.uint8, device=dev) if want_s8 else None
    s64 = torch.empty(ncells, d, d, d, dtype=torch.float64, device=dev) if want_s64 else None
    _lib.call("icsg3d_voxelize", _p(sites), _p(nsites), _p(lattice), ncells, max_sites, d, ctypes.c_double(eps_frac),
              _p(m32), _p(m64), _p(s8), _p(s64), _stream())
    return m32, m64, s8, s64


def density_matrix(N, z, l, dims=(32, 32, 32), sigma=0.5, dist=None, label_frac=1.0, eps_frac=0.25, device="cuda"):
    """Reference signature (utils.py:97-100; `dist` is accepted and ignored exactly like the reference, which
    always calls scipy's euclidean cdist).  Returns numpy float64 (M, S) of shape `dims`."""
    if not (dims[0] == dims[1] == dims[2]):
        raise ValueError("icsg3d voxeliser: cubic grids only (the reference is hard-wired to d=32)")
    z = np.asarray(z, dtype=np.float64)
    if len(z) > MAX_SITES:
        raise ValueError(f"more than {MAX_SITES} sites (the reference skips cells above max_sites=40)")
    rec = site_records(N, z, sigma, label_frac)[None]
    sites = torch.from_numpy(rec).to(device)
    nsites = torch.tensor([len(z)], dtype=torch.int32, device=device)
    lat = torch.tensor(np.asarray(l, dtype=np.float64)[:3].reshape(1, 3), dtype=torch.float64, device=device)
    _, m64, _, s64 = voxelize_cells(sites, nsites, lat, d=dims[0], eps_frac=eps_frac, want_m32=False, want_m64=True,
                                    want_s8=False, want_s64=True)
    return m64[0].cpu().numpy(), s64[0].cpu().numpy()


def coordinate_grid(l, dim=32, eps_frac=0.25, device="cuda"):
    """utils.py:88-94 — channels 1..3 of the voxeliser's fp32 network input, returned as (dim,dim,dim,3)."""
    rec = site_records(np.zeros((1, 3)), np.ones(1), np.ones(1))[None]
    sites = torch.from_numpy(rec).to(device)
    nsites = torch.tensor([1], dtype=torch.int32, device=device)
    lat = torch.tensor(np.asarray(l, dtype=np.float64)[:3].reshape(1, 3), dtype=torch.float64, device=device)
    m32, _, _, _ = voxelize_cells(sites, nsites, lat, d=dim, eps_frac=eps_frac, want_s8=False)
    return m32[0, ..., 1:].double().cpu().numpy()


def synthetic_cells(ncells, seed=0, label_frac=1.0, device="cuda", max_sites=8):
    """On-device perovskite-like ABX3 cells: returns (sites, nsites, lattice) device tensors."""
    sites = torch.zeros(ncells, max_sites, _REC, dtype=torch.float64, device=device)
    nsites = torch.zeros(ncells, dtype=torch.int32, device=device)
    lat = torch.zeros(ncells, 3, dtype=torch.float64, device=device)
    _lib.call("icsg3d_synth_perovskite_sites", ctypes.c_uint64(seed), ncells, max_sites, ctypes.c_double(label_frac),
              _p(sites), _p(nsites), _p(lat), _stream())
    return sites, nsites, lat


def synthetic_batch(n, d=32, seed=0, ncond=10, device="cuda"):
    """Synthetic training batch generated entirely on the device: M fp32 (n,d,d,d,4), one-hot cond (n,ncond),
    species uint8 (n,d,d,d)."""
    sites, nsites, lat = synthetic_cells(n, seed=seed, device=device)
    m32, _, s8, _ = voxelize_cells(sites, nsites, lat, d=d)
    g = torch.Generator(device=device).manual_seed(seed)
    bins = torch.randint(0, ncond, (n,), generator=g, device=device)
    cond = torch.nn.functional.one_hot(bins, ncond).to(torch.float32)
    return m32, cond, s8


def to_lattice_params(p, eps_frac=0.25, d=32, axis=(-3, -2, -1)):
    """utils.py:160-178 on a (B,d,d,d,3) numpy array or torch tensor (device tensors stay on the device)."""
    if torch.is_tensor(p):
        mx = p.amax(dim=(1, 2, 3))
        mn = p.amin(dim=(1, 2, 3))
        ap = (mx - mn) / (1 + 2 * eps_frac) / (1 - 1.0 / d)
        return ap - ap / d
    mx = p.max(axis=(1, 2, 3))
    mn = p.min(axis=(1, 2, 3))
    ap = (mx - mn) / (1 + 2 * eps_frac) / (1 - 1.0 / d)
    return ap - ap / d


def to_voxel_params(lp, eps=0.25, d=32):
    """utils.py:181-190."""
    return (lp + (2 * lp * eps)) / d
