"""Exhaustive timing of the halo kernel's (TD, TH, NT) configurations on the layer shapes it serves in the VAE+DFC step
(batch 32), against the planner's own choice and the per-tap kernel.  usage: python tools/halo_autotune.py [B]"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = _lib.lib()
SHAPES = {"c4": (16, 64, 128), "c4d": (16, 128, 64), "c5": (8, 128, 128), "c6": (8, 128, 256), "c6d": (8, 256, 128),
          "dec2": (8, 128, 64), "dec2d": (8, 64, 128), "enc3x3": (8, 96, 64)}


def plan(D, cin, cout):
    out = (ctypes.c_int * 10)()
    L.icsg3d_conv3d_k3_plan(B, D, D, D, cin, cout, 148, out)
    return list(out)


def timeit(x, wp, bias, y, ws, iters=20):
    for _ in range(3):
        ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y, ws=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y, ws=ws)
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
    ws = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _lib.call("icsg3d_conv3d_halo_force", 0, 0, 0)
    _lib.call("icsg3d_conv3d_set_impl", 0)
    p0 = plan(D, cin, cout)
    t_auto = timeit(x, wp, bias, y, ws)
    _lib.call("icsg3d_conv3d_set_impl", 1)
    t_v1 = timeit(x, wp, bias, y, ws)
    _lib.call("icsg3d_conv3d_set_impl", 2)   # force halo: the dispatch rules do not filter the forced configurations
    rows = []
    for nt in sorted({n for n in (cout, cout // 2, cout // 4, 128, 64) if 16 <= n <= 256 and cout % n == 0}, reverse=True):
        TH = D
        while TH >= 2:
            for TD in range(1, min(8, D) + 1):
                _lib.call("icsg3d_conv3d_halo_force", TD, TH, nt)
                p = plan(D, cin, cout)
                if p[0] != 1 or p[1] != TD or p[2] != TH or p[4] != nt:
                    continue
                rows.append((timeit(x, wp, bias, y, ws, 10), TD, TH, nt, p[3], p[5], p[6], p[7]))
            TH //= 2
    _lib.call("icsg3d_conv3d_halo_force", 0, 0, 0)
    _lib.call("icsg3d_conv3d_set_impl", 0)
    rows.sort()
    gf = 2.0 * B * D ** 3 * 27 * cin * cout / 1e9
    res[name] = {"shape": (B, D, cin, cout), "planner": p0[:8], "us_planner": t_auto, "us_pertap": t_v1,
                 "best": [dict(us=r[0], TD=r[1], TH=r[2], NT=r[3], G=r[4], a_bufs=r[5], b_stages=r[6], items=r[7]) for r in rows[:5]]}
    print(f"{name:7s} {B}x{D}^3 {cin}->{cout}: planner {p0[1:8]} {t_auto:.1f} us ({gf / t_auto / 1e-3 / 1e3:.0f} TF/s) | per-tap {t_v1:.1f} | best "
          + ", ".join(f"TD{r[1]} TH{r[2]} NT{r[3]} G{r[4]} ab{r[5]} bs{r[6]}: {r[0]:.1f}" for r in rows[:4]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/halo_autotune.json", "w"), indent=1)
