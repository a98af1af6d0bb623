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


def _as_dev(p, device):
    """numpy / torch array -> contiguous device tensor in fp32 (kept) or fp64 (anything else floating)."""
    t = p if torch.is_tensor(p) else torch.from_numpy(np.ascontiguousarray(p))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t.to(device).contiguous()


def lattice_params_device(p, c0=0, eps_frac=0.25, d=None, x16=None, want_dv=True):
    """Device form of to_lattice_params + to_voxel_params (utils.py:160-190): p is a CUDA tensor (B,d,d,d,ld), fp32 or
    fp64, whose channels [c0, c0+3) are the coordinate grids; returns device tensors (lp (B,3), dv (B,3)) in p's dtype.
    `x16` (bf16 (B,d,d,d,16), only for the fp32 4-channel decoder output with c0 == 1) is filled with the packed U-Net
    input in the same pass (generate.py:208-220: the decoder output is read once)."""
    if not p.is_cuda:
        raise _lib.Icsg3dError("lattice_params_device needs a CUDA tensor (no CPU fallback)")
    B, ld = p.shape[0], p.shape[-1]
    d = d or p.shape[1]
    vox = p.numel() // (B * ld)
    dt = 1 if p.dtype == torch.float32 else 2
    nsplit = int(_lib.lib().icsg3d_lattice_nsplit(B, ctypes.c_int64(vox)))
    part = torch.empty(B, nsplit, 6, dtype=p.dtype, device=p.device)
    lp = torch.empty(B, 3, dtype=p.dtype, device=p.device)
    dv = torch.empty(B, 3, dtype=p.dtype, device=p.device) if want_dv else None
    _lib.call("icsg3d_coord_minmax", _p(p), dt, ld, c0, B, ctypes.c_int64(vox), nsplit, _p(part), _p(x16), _stream())
    _lib.call("icsg3d_lattice_finalize", _p(part), dt, B, nsplit, ctypes.c_double(eps_frac), d, _p(lp), _p(dv), _stream())
    return lp, dv


def to_lattice_params(p, eps_frac=0.25, d=32, axis=(-3, -2, -1), device="cuda"):
    """utils.py:160-178 on a (B,d,d,d,3) array: numpy in -> numpy out (same dtype rules as the reference: float32 stays
    float32), CUDA tensor in -> CUDA tensor out.  The min/max reduction and the arithmetic (incl. the a*(1-1/d) quirk)
    run in libicsg3d (csrc/post.cu)."""
    if torch.is_tensor(p) and p.is_cuda:
        return lattice_params_device(p, 0, eps_frac, d, want_dv=False)[0]
    lp, _ = lattice_params_device(_as_dev(p, device), 0, eps_frac, d, want_dv=False)
    return lp.cpu().numpy()


def to_voxel_params(lp, eps=0.25, d=32):
    """utils.py:181-190 (three multiply-adds per sample on (B,3) lattice parameters: host arithmetic; the device path
    gets dv from lattice_params_device in the same launch as lp)."""
    return (lp + (2 * lp * eps)) / d


# ---- random_rotation_3d (utils.py:193-222): exact signed axis permutations on the device --------------------
ROT_AXES = [(0, 1), (0, 2), (1, 2)]


def rot90_transform(axes_seq):
    """Compose scipy.ndimage.rotate(., 90, axes=(a,b), reshape=False) == np.rot90(., 1, axes=(a,b)) over a sequence of
    axis pairs into one signed permutation (perm[3], flip[3]): out[o] = in[s], s_x = flip_x ? d-1-o[perm_x] : o[perm_x]."""
    perm, flip = [0, 1, 2], [0, 0, 0]
    for a, b in axes_seq:
        p1, f1 = [0, 1, 2], [0, 0, 0]
        p1[a], p1[b], f1[b] = b, a, 1          # one rotation: s_a = o_b, s_b = d-1-o_a
        # new total = previous total applied to the result of this rotation:  s = s_prev(s_1(o))
        perm, flip = [p1[perm[x]] for x in range(3)], [flip[x] ^ f1[perm[x]] for x in range(3)]
    return perm, flip


def rotate90_batch(x, xforms):
    """x: CUDA tensor (B,d,d,d[,C]) of any dtype; xforms: list of (perm, flip) per sample (or one for all)."""
    if not x.is_cuda:
        raise _lib.Icsg3dError("rotate90_batch needs a CUDA tensor (no CPU fallback)")
    x = x.contiguous()
    B, d = x.shape[0], x.shape[1]
    if isinstance(xforms, tuple):
        xforms = [xforms] * B
    tf = torch.tensor([list(p) + list(f) for p, f in xforms], dtype=torch.int32).to(x.device)
    vb = x.element_size() * (x.numel() // (B * d ** 3))
    out = torch.empty_like(x)
    _lib.call("icsg3d_rotate90_batch", _p(x), _p(out), B, d, vb, _p(tf), _stream())
    return out


def random_rotation_3d(M, S, p, rot_angle=90, nrotations=3, device="cuda"):
    """Drop-in for utils.py:193-222 (same np.random draw, same return order/dtypes).  The reference rotates by exactly 90
    degrees with cubic-spline interpolation, i.e. a signed axis permutation up to ~5e-16 of spline round-off; the kernel
    applies the composed permutation exactly: S is bit-identical, M / p agree to that round-off."""
    if rot_angle != 90:
        raise NotImplementedError("only the reference's own rot_angle=90 (exact axis permutations) is implemented")
    rotations = [ROT_AXES[x] for x in np.random.choice(3, 3)]
    xf = rot90_transform(rotations[:nrotations])
    outs = []
    for a in (M, S, p):
        t = torch.from_numpy(np.ascontiguousarray(a))[None].to(device)
        outs.append(rotate90_batch(t, xf)[0].cpu().numpy())
    M_rot, S_rot, p_rot = outs
    S_rot = np.abs(np.rint(S_rot))
    p_rot[np.abs(p_rot) < 1e-14] = 0
    assert np.array_equal(np.unique(S_rot), np.unique(S))
    return M_rot, S_rot, p_rot
