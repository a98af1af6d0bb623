"""Shared helpers for the parity tests."""
import numpy as np
import torch


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().flatten().cpu()
    b = torch.as_tensor(b).detach().double().flatten().cpu()
    return float(torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30))


def cosine(a, b):
    a = torch.as_tensor(a).detach().double().flatten().cpu()
    b = torch.as_tensor(b).detach().double().flatten().cpu()
    return float((a @ b) / (torch.linalg.norm(a) * torch.linalg.norm(b) + 1e-30))


def synthetic_batch(B, d=32, seed=0, ncond=10):
    """Voxelised perovskite-like cells via the ORACLE voxeliser (numpy): M (B,d,d,d,4) fp32, cond one-hot, species."""
    from oracle import voxelizer as vox
    rng = np.random.default_rng(seed)
    M = np.zeros((B, d, d, d, 4), dtype=np.float32)
    S = np.zeros((B, d, d, d), dtype=np.uint8)
    for b in range(B):
        N, z, l, sigma = vox.synthetic_cell(rng)
        dens, spec = vox.density_matrix(N, z, l, dims=(d, d, d), sigma=sigma)
        M[b, ..., 0] = dens
        M[b, ..., 1:] = vox.coordinate_grid(l, dim=d)
        S[b] = spec.astype(np.uint8)
    cond = np.eye(ncond, dtype=np.float32)[rng.integers(0, ncond, size=B)]
    return torch.from_numpy(M), torch.from_numpy(cond), torch.from_numpy(S)
