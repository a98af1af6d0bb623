"""Time the VAE+DFC layer shapes (fprop and dgrad operand shapes, B=32) with each conv kernel generation in ONE
process.  usage: conv_sweep.py [impls=0,2] [iters=10]   (0 auto/stream, 1 per-tap, 2 halo)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib, ops

impls = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,2").split(",")]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
SHAPES = [("c1.f", 32, 32, 16, 32), ("c1.d", 32, 32, 32, 16), ("c2.f", 32, 32, 32, 64), ("c2.d", 32, 32, 64, 32),
          ("enc1.f/out", 32, 32, 16, 16), ("dec4.f", 32, 32, 32, 16), ("c3", 32, 16, 64, 64), ("c4.f", 32, 16, 64, 128),
          ("c4.d", 32, 16, 128, 64), ("dec3.f", 32, 16, 64, 32), ("dec3.d", 32, 16, 32, 64), ("enc2.f", 32, 16, 16, 32),
          ("c5", 32, 8, 128, 128), ("c6.f", 32, 8, 128, 256), ("c6.d", 32, 8, 256, 128), ("c9.f", 32, 4, 256, 512),
          ("c9.d", 32, 4, 512, 256), ("c10", 32, 4, 512, 512), ("enc4.f", 32, 4, 64, 128), ("enc4.d", 32, 4, 128, 64),
          ("enc5.f", 32, 2, 128, 16), ("dec1.f", 32, 4, 16, 128), ("dec2.f", 32, 8, 128, 64)]
res = []
for name, B, D, cin, cout in SHAPES:
    x = torch.randn(B, D, D, D, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5
    wp = ops.pack_conv_w_fprop(w)
    bias = torch.zeros(cout, device="cuda")
    y = torch.empty(B, D, D, D, cout, dtype=torch.bfloat16, device="cuda")
    ws = torch.empty(max(ops.conv3d_k3_workspace_bytes(B, D, cin, cout), 16), dtype=torch.uint8, device="cuda")
    row = {"layer": name, "B": B, "D": D, "cin": cin, "cout": cout}
    for impl in impls:
        _lib.call("icsg3d_conv3d_set_impl", impl)
        for _ in range(2):
            ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y, ws=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y, ws=ws)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        row[f"impl{impl}_us"] = us
        row[f"impl{impl}_tflops"] = 2.0 * B * D ** 3 * 27 * cin * cout / us / 1e6
    _lib.call("icsg3d_conv3d_set_impl", 0)
    res.append(row)
    print(" ".join(f"{k}={v:.1f}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items()), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/conv_sweep.json", "w"), indent=1)
