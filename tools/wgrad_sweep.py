"""Time the filter-gradient layer shapes (B=32) with the plane-streaming kernel (impl 0) and the per-tap kernel (impl 1)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib, ops

SHAPES = [("dec4", 32, 32, 32, 16), ("enc1/out", 32, 32, 16, 16), ("dec3", 32, 16, 64, 32), ("enc2", 32, 16, 16, 32),
          ("enc3", 32, 8, 32, 64), ("dec2", 32, 8, 128, 64), ("enc4", 32, 4, 64, 128)]
res = []
for name, B, D, cin, cout in SHAPES:
    x = torch.randn(B, D, D, D, cin, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, D, D, D, cout, device="cuda").to(torch.bfloat16)
    out = torch.empty(27, cin, cout, device="cuda")
    row = {"layer": name, "B": B, "D": D, "cin": cin, "cout": cout}
    for impl in (0, 1):
        _lib.call("icsg3d_conv3d_set_impl", impl)
        for _ in range(2):
            ops.conv3d_k3_wgrad(x, dy, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv3d_k3_wgrad(x, dy, out=out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        row[f"impl{impl}_us"] = us
        row[f"impl{impl}_tflops"] = 2.0 * B * D ** 3 * 27 * cin * cout / us / 1e6
    _lib.call("icsg3d_conv3d_set_impl", 0)
    res.append(row)
    print(" ".join(f"{k}={v:.1f}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items()), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/wgrad_sweep.json", "w"), indent=1)
