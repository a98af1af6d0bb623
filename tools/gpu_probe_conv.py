"""GPU bring-up probe for the tcgen05 Conv3D kernels (run under gpurun).

    python tools/gpu_probe_conv.py            # driver: every case in its own subprocess + timeout
    python tools/gpu_probe_conv.py CASE_JSON  # worker: one case

Each case compares the tcgen05 kernel against (a) the CUDA-core cross-check kernel on the device and
(b) torch's CPU conv3d (fp32, on the same bf16-rounded operands) for the small shapes.
Writes one JSON line per case to gpurun_out/conv_probe.jsonl.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = []
for B, D, cin, cout in [(2, 32, 16, 16), (2, 32, 32, 64), (2, 32, 64, 32), (2, 16, 64, 128), (4, 8, 128, 256),
                        (8, 4, 256, 512), (8, 4, 512, 512), (16, 2, 128, 16), (4, 2, 128, 16), (2, 32, 16, 32),
                        (1, 64, 32, 64)]:
    CASES.append({"kind": "fprop", "B": B, "D": D, "cin": cin, "cout": cout})
for B, D, cin, cout in [(2, 32, 16, 16), (2, 32, 32, 16), (2, 16, 64, 32), (4, 8, 128, 64), (8, 4, 64, 128),
                        (16, 2, 128, 16), (8, 4, 16, 128), (2, 16, 32, 64)]:
    CASES.append({"kind": "wgrad", "B": B, "D": D, "cin": cin, "cout": cout})
for kc in (64, 32, 16):
    CASES.append({"kind": "shift", "kc": kc})


def rel_l2(a, b):
    import torch
    a = a.double().flatten()
    b = b.double().flatten()
    return float(torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30))


def worker(case):
    import torch
    from icsg3d_b200 import ops, _lib
    import ctypes

    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    res = dict(case)
    if case["kind"] == "fprop":
        B, D, cin, cout = case["B"], case["D"], case["cin"], case["cout"]
        x = torch.randn(B, D, D, D, cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(3, 3, 3, cin, cout, device=dev) / (27 * cin) ** 0.5).float()
        bias = torch.randn(cout, device=dev).float()
        wp = ops.pack_conv_w_fprop(w)
        t0 = time.time()
        y = ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU)
        torch.cuda.synchronize()
        res["t_first_ms"] = (time.time() - t0) * 1e3
        yr = ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, ref=True)
        torch.cuda.synchronize()
        res["rel_l2_vs_ref"] = rel_l2(y.float(), yr)
        res["max_abs_vs_ref"] = float((y.float() - yr).abs().max())
        # fp32 output path + no activation
        y32 = ops.conv3d_k3(x, wp, None, act=ops.ACT_NONE, out_dtype=torch.float32)
        yr32 = ops.conv3d_k3(x, wp, None, act=ops.ACT_NONE, ref=True)
        torch.cuda.synchronize()
        res["rel_l2_f32_vs_ref"] = rel_l2(y32, yr32)
        if B * D ** 3 * cin * cout <= 2 * 32 ** 3 * 64 * 64:
            xc = x.float().cpu().permute(0, 4, 1, 2, 3)
            wc = w.to(torch.bfloat16).float().cpu().permute(4, 3, 0, 1, 2)
            yc = torch.nn.functional.conv3d(xc, wc, None, padding=1).permute(0, 2, 3, 4, 1)
            res["rel_l2_f32_vs_torch_cpu"] = rel_l2(y32.cpu(), yc)
        # timing
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y)
        ev0.record()
        n = 10
        for _ in range(n):
            ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / n
        res["ms"] = ms
        res["tflops"] = 2.0 * B * D ** 3 * 27 * cin * cout / ms / 1e9
    elif case["kind"] == "wgrad":
        B, D, cin, cout = case["B"], case["D"], case["cin"], case["cout"]
        x = torch.randn(B, D, D, D, cin, device=dev).to(torch.bfloat16)
        dy = torch.randn(B, D, D, D, cout, device=dev).to(torch.bfloat16)
        dw = ops.conv3d_k3_wgrad(x, dy)
        dwr = ops.conv3d_k3_wgrad(x, dy, ref=True)
        torch.cuda.synchronize()
        res["rel_l2_vs_ref"] = rel_l2(dw, dwr)
        res["max_abs_vs_ref"] = float((dw - dwr).abs().max())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.conv3d_k3_wgrad(x, dy, out=dw)
        ev0.record()
        n = 10
        for _ in range(n):
            ops.conv3d_k3_wgrad(x, dy, out=dw)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / n
        res["ms"] = ms
        res["tflops"] = 2.0 * B * D ** 3 * 27 * cin * cout / ms / 1e9
    elif case["kind"] == "shift":
        kc, n, nshift, rows = case["kc"], 64, 16, 160
        a = torch.randn(rows, kc, device=dev).to(torch.bfloat16)
        b = torch.randn(n, kc, device=dev).to(torch.bfloat16)
        out = torch.zeros(2, nshift, 128, n, device=dev)
        _lib.call("icsg3d_probe_shifted_desc", ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                  ctypes.c_void_p(out.data_ptr()), rows, kc, n, nshift,
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        for mode in range(2):
            errs = []
            for s in range(nshift):
                exp = a[s:s + 128].float() @ b.float().t()
                errs.append(round(rel_l2(out[mode, s], exp), 5))
            res[f"mode{mode}_rel_l2_by_shift"] = errs
    return res


def main():
    if len(sys.argv) > 1:
        case = json.loads(sys.argv[1])
        try:
            res = worker(case)
        except Exception as e:  # noqa: BLE001
            res = dict(case)
            res["error"] = f"{type(e).__name__}: {e}"
        print("RESULT " + json.dumps(res))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "conv_probe.jsonl"), "w")
    for case in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, json.dumps(case)], capture_output=True, text=True,
                               timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                rec = json.loads(line[-1][7:])
            else:
                rec = dict(case)
                rec["error"] = "no result"
            if r.returncode != 0 or "error" in rec:
                rec["stderr_tail"] = r.stderr[-600:]
                rec["stdout_tail"] = r.stdout[-400:]
        except subprocess.TimeoutExpired:
            rec = dict(case)
            rec["error"] = "timeout"
        out.write(json.dumps(rec) + "\n")
        out.flush()
        print(json.dumps(rec))


if __name__ == "__main__":
    main()
