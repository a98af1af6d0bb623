"""Timing of the plane-streaming conv kernel for every number of h-blocks per plane on the layer shapes it serves in the
VAE+DFC step (batch 32) against the planner's own choice.  usage: python tools/stream_autotune.py [B]"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = _lib.lib()
SHAPES = {"c1": (32, 16, 32), "c1d": (32, 32, 16), "c2": (32, 32, 64), "c2d": (32, 64, 32), "c3": (16, 64, 64), "dec3": (16, 64, 32),
          "dec3d": (16, 32, 64), "dec4": (32, 32, 16), "dec4d": (32, 16, 32), "dout": (32, 16, 16), "enc2x3": (16, 48, 32)}


def plan(D, cin, cout):
    out = (ctypes.c_int * 10)()
    L.icsg3d_conv3d_k3_plan(B, D, D, D, cin, cout, 148, out)
    return list(out)


def timeit(x, wp, bias, y, iters=20):
    for _ in range(3):
        ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


res = {}
for name, (D, cin, cout) in SHAPES.items():
    x = torch.randn(B, D, D, D, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5
    wp = ops.pack_conv_w_fprop(w)
    bias = torch.zeros(cout, device="cuda")
    y = torch.empty(B, D, D, D, cout, dtype=torch.bfloat16, device="cuda")
    _lib.call("icsg3d_conv3d_stream_force", 0)
    p0 = plan(D, cin, cout)
    t_auto = timeit(x, wp, bias, y)
    rows = []
    for nb in range(1, D + 1):
        _lib.call("icsg3d_conv3d_stream_force", nb)
        p = plan(D, cin, cout)
        if p[0] != 2:
            continue
        rows.append((timeit(x, wp, bias, y, 10), nb, p[1:7]))
    _lib.call("icsg3d_conv3d_stream_force", 0)
    rows.sort()
    res[name] = {"shape": (B, D, cin, cout), "planner": p0[:8], "us_planner": t_auto,
                 "best": [dict(us=r[0], n_hblk=r[1], plan_R_TH_T_C_stages_issuers=r[2]) for r in rows[:5]]}
    print(f"{name:7s} {B}x{D}^3 {cin}->{cout}: planner {p0[1:7]} {t_auto:.1f} us | best "
          + ", ".join(f"nb{r[1]} {r[2]}: {r[0]:.1f}" for r in rows[:4]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/stream_autotune.json", "w"), indent=1)
