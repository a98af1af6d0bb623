"""Role timeline of the plane-streaming conv kernel (CTA 1): where the producer, the first MMA issuer and the two
epilogue warps of lane quarter 0 spend each plane step.  usage: stream_timeline.py [steps=24]
Output: gpurun_out/stream_timeline.json + a per-step table (cycles, relative to the issuer's first stamp)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib, ops

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
SHAPES = [("c1.f", 32, 32, 16, 32), ("c2.f", 32, 32, 32, 64), ("out.f", 32, 32, 16, 16), ("c3", 32, 16, 64, 64)]
out = {}
for name, B, D, cin, cout in SHAPES:
    x = torch.randn(B, D, D, D, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5
    wp = ops.pack_conv_w_fprop(w)
    bias = torch.zeros(cout, device="cuda")
    y = torch.empty(B, D, D, D, cout, dtype=torch.bfloat16, device="cuda")
    ws = torch.empty(max(ops.conv3d_k3_workspace_bytes(B, D, cin, cout), 16), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y, ws=ws)
    buf = torch.zeros(4, steps, 4, dtype=torch.int64, device="cuda")
    _lib.call("icsg3d_conv3d_stream_debug", buf.data_ptr(), steps)
    ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y, ws=ws)
    torch.cuda.synchronize()
    _lib.call("icsg3d_conv3d_stream_debug", None, 0)
    t = buf.cpu()
    t0 = int(t[1, 0, 0])
    rel = torch.where(t > 0, t - t0, torch.zeros_like(t))
    out[name] = rel.tolist()
    print(f"== {name} (B={B}, {D}^3, {cin}->{cout}); cycles relative to the issuer's first stamp")
    print("step | producer TMA | issuer: enter slot_free plane_landed mmas_issued | epi0: enter acc_done drained released | epi1: ...")
    for s in range(steps):
        r = rel[:, s, :].tolist()
        print(f"{s:3d} | {r[0][0]:8d} | " + " ".join(f"{v:8d}" for v in r[1]) + " | " + " ".join(f"{v:8d}" for v in r[2]) +
              " | " + " ".join(f"{v:8d}" for v in r[3]))
    # steady-state periods (median of the deltas between consecutive steps, middle of the segment)
    def period(role, col):
        v = [int(rel[role, s + 1, col] - rel[role, s, col]) for s in range(4, steps - 2) if rel[role, s + 1, col] > 0 and rel[role, s, col] > 0]
        v.sort()
        return v[len(v) // 2] if v else 0
    print("median period: issuer", period(1, 3), "epi0", period(2, 3), "epi1", period(3, 3),
          "| issuer wait slot", sorted(int(rel[1, s, 1] - rel[1, s, 0]) for s in range(4, steps - 2))[(steps - 6) // 2],
          "wait plane", sorted(int(rel[1, s, 2] - rel[1, s, 1]) for s in range(4, steps - 2))[(steps - 6) // 2],
          "issue", sorted(int(rel[1, s, 3] - rel[1, s, 2]) for s in range(4, steps - 2))[(steps - 6) // 2],
          "| epi0 wait acc", sorted(int(rel[2, s, 1] - rel[2, s, 0]) for s in range(4, steps - 2))[(steps - 6) // 2],
          "drain", sorted(int(rel[2, s, 2] - rel[2, s, 1]) for s in range(4, steps - 2))[(steps - 6) // 2],
          "release", sorted(int(rel[2, s, 3] - rel[2, s, 2]) for s in range(4, steps - 2))[(steps - 6) // 2], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/stream_timeline.json", "w"))
