"""Probe 4 driver: MN-major UMMA operands with overlapping MN blocks (wgrad tap folding).  Prints rel-L2 per config."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib

res = []
for (ca, a_shift, cb, nblk_b, b_shift, ksteps) in [
        (32, 8, 16, 1, 8, 1), (32, 8, 16, 3, 8, 2), (32, 1, 16, 1, 8, 2), (32, 8, 16, 3, 33, 2), (32, 1, 16, 3, 33, 4),
        (16, 1, 16, 3, 33, 4), (16, 1, 16, 3, 17, 4), (64, 1, 16, 3, 33, 4), (64, 1, 32, 3, 33, 4), (32, 1, 32, 3, 17, 4),
        (32, 1, 64, 3, 17, 2), (16, 3, 16, 3, 1, 2), (32, 33, 16, 3, 1, 4)]:
    rows = 256
    if 16 * ksteps + (128 // ca - 1) * a_shift > rows or 16 * ksteps + (nblk_b - 1) * b_shift > rows:
        continue
    g = torch.Generator().manual_seed(ca * 1000 + cb + a_shift)
    X = torch.randn(rows, ca, generator=g).to(torch.bfloat16)
    Y = torch.randn(rows, cb, generator=g).to(torch.bfloat16)
    n = nblk_b * cb
    out = torch.zeros(128, n, device="cuda")
    Xd, Yd = X.cuda(), Y.cuda()
    _lib.call("icsg3d_probe_mn_fold", ctypes.c_void_p(Xd.data_ptr()), ctypes.c_void_p(Yd.data_ptr()),
              ctypes.c_void_p(out.data_ptr()), rows, ca, cb, a_shift, nblk_b, b_shift, ksteps,
              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    K = 16 * ksteps
    ref = torch.zeros(128, n)
    Xf, Yf = X.float(), Y.float()
    for j in range(128 // ca):
        for l in range(nblk_b):
            ref[j * ca:(j + 1) * ca, l * cb:(l + 1) * cb] = Xf[j * a_shift:j * a_shift + K].T @ Yf[l * b_shift:l * b_shift + K]
    err = float((out.cpu() - ref).norm() / ref.norm())
    row = dict(ca=ca, a_shift=a_shift, cb=cb, nblk_b=nblk_b, b_shift=b_shift, ksteps=ksteps, rel_l2=err)
    res.append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mnfold_probe.json", "w"), indent=1)
