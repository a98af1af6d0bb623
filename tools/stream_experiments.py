"""TIMING-ONLY experiments on the plane-streaming conv kernel: which epilogue TMEM operation costs the MMA pipe time.
Runs tools/conv_sweep.py-style timings of c1/c2/c3 with ICSG3D_STREAM_EXP = 0, 1 (no slot zeroing), 2 (no TMEM loads),
3, 4 (no global stores), 7 in separate processes (the results of the flagged runs are wrong by construction)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from icsg3d_b200 import ops
out = {}
for name, B, D, cin, cout in [("c1.f", 32, 32, 16, 32), ("c2.f", 32, 32, 32, 64), ("c2.d", 32, 32, 64, 32), ("c3", 32, 16, 64, 64), ("out.f", 32, 32, 16, 16)]:
    x = torch.randn(B, D, D, D, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5
    wp = ops.pack_conv_w_fprop(w)
    bias = torch.zeros(cout, device="cuda")
    y = torch.empty(B, D, D, D, cout, dtype=torch.bfloat16, device="cuda")
    ws = torch.empty(max(ops.conv3d_k3_workspace_bytes(B, D, cin, cout), 16), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y, ws=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y, ws=ws)
    e1.record()
    torch.cuda.synchronize()
    out[name] = e0.elapsed_time(e1) / 20 * 1e3
print(json.dumps(out))
''' % ROOT
res = {}
variants = [(f, 0) for f in (0, 1, 2, 4, 7)] + [(0, n) for n in (1, 2, 3)] + [(7, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "issuers":
    variants = [(0, n) for n in (1, 2, 3)] + [(7, 1)]
for flags, issuers in variants:
    env = dict(os.environ, ICSG3D_STREAM_EXP=str(flags))
    if issuers:
        env["ICSG3D_STREAM_ISSUERS"] = str(issuers)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    line = r.stdout.strip().split("\n")[-1] if r.stdout.strip() else ""
    try:
        res[f"exp{flags}_issuers{issuers}"] = json.loads(line)
    except Exception:  # noqa: BLE001
        res[f"exp{flags}_issuers{issuers}"] = {"error": (r.stderr or "")[-400:]}
    print("exp_flags", flags, "issuers", issuers or "default", res[f"exp{flags}_issuers{issuers}"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "stream_experiments.json"), "w"), indent=1)
