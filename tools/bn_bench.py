"""HBM roofline of the BatchNorm passes at the VAE+DFC step's main shapes (B=32): algorithmic bytes / time vs the measured
copy bandwidth (MEASURED_PEAKS.json).  usage: bn_bench.py [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import ops
from icsg3d_b200.ops import ACT_LEAKY, ACT_NONE, POST_NONE, POST_POOL2, POST_UP2

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    PEAK = 6650.0
BF = torch.bfloat16
dev = "cuda"
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MB, only ever READ: evicts L2 with clean lines


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.sum()  # evict L2 (126 MB) with clean lines so that every pass streams from HBM as inside the step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


res = []
# (name, D, C, act, post, pre_relu, tap)
CASES = [("pm.c2 (64ch@32^3, pool, tap)", 32, 64, ACT_NONE, POST_POOL2, True, True),
         ("pm.c1 (32ch@32^3)", 32, 32, ACT_NONE, POST_NONE, True, False),
         ("pm.c4 (128ch@16^3, pool, tap)", 16, 128, ACT_NONE, POST_POOL2, True, True),
         ("enc1 (16ch@32^3, leaky, pool)", 32, 16, ACT_LEAKY, POST_POOL2, False, False),
         ("dec4 (16ch@32^3, leaky)", 32, 16, ACT_LEAKY, POST_NONE, False, False),
         ("dec3 (32ch@16^3, leaky, up2)", 16, 32, ACT_LEAKY, POST_UP2, False, False)]
B = 32
for name, D, C, act, post, pre_relu, tap in CASES:
    x = torch.randn(B, D, D, D, C, device=dev).to(BF)
    other = torch.randn(B, D, D, D, C, device=dev).to(BF) if tap else None
    Do = D // 2 if post == POST_POOL2 else (2 * D if post == POST_UP2 else D)
    y = torch.empty(B, Do, Do, Do, C, dtype=BF, device=dev)
    idx = torch.empty(B, Do, Do, Do, C, dtype=torch.uint8, device=dev) if post == POST_POOL2 else None
    dy = torch.randn(B, Do, Do, Do, C, device=dev).to(BF)
    dx = torch.empty_like(x)
    mean, rstd, scale, shift = (torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.ones(C, device=dev),
                                torch.zeros(C, device=dev))
    rows = x.numel() // C
    n = ops.bn_nparts(rows, C, BF)
    part = torch.zeros(n, 2, C, dtype=torch.float64, device=dev)
    nb = ops.bn_bwd_nparts(x, C, post)
    partb = torch.zeros(nb, 2, C, dtype=torch.float64, device=dev)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    xb, yb, ib = x.numel() * 2, y.numel() * 2, (idx.numel() if idx is not None else 0)
    passes = [
        ("stats", lambda: ops.bn_stats(x, C, part), xb),
        ("apply_fwd", lambda: ops.bn_apply_fwd(x, C, scale, shift, act, post, y=y, pool_idx=idx), xb + yb + ib),
        ("bwd_reduce", lambda: ops.bn_bwd_reduce(dy, x, C, mean, rstd, scale, shift, act, post, idx, partb), xb + yb + ib),
        ("bwd_apply", lambda: ops.bn_bwd_apply(dy, x, C, mean, rstd, scale, shift, act, post, idx, sums, float(rows), dx,
                                               pre_relu=pre_relu, tap_other=other, tap_coef=0.1),
         xb + yb + ib + xb + (xb if tap else 0)),
    ]
    for pname, fn, nbytes in passes:
        us = timeit(fn)
        row = {"layer": name, "pass": pname, "us": us, "MB": nbytes / 1e6, "GBps": nbytes / us / 1e3, "frac_hbm": nbytes / us / 1e3 / PEAK}
        res.append(row)
        print(f"{name:34s} {pname:11s} {us:8.1f} us {nbytes / 1e6:8.1f} MB {row['GBps']:8.0f} GB/s {100 * row['frac_hbm']:5.1f}%", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bn_bench.json", "w"), indent=1)
