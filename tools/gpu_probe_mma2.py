"""MMA rate probe at the streaming kernel's operand shapes: swizzle 32/64/128 B, N = 96/192 (kd folded), operand start
addresses advancing by 0 / one row / one padded h-row per MMA (tap shifts)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icsg3d_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
res = []
for swz in (32, 64, 128):
    for n in (32, 64, 96, 128, 192):
        for (a_step, b_step) in ((0, 0), (swz, 0), (33 * swz, 0), (0, n * swz), (swz, n * swz), (33 * swz, n * swz)):
            if 7 * b_step + n * swz > 48 * 1024 or 7 * a_step + 128 * swz > 48 * 1024:
                continue
            reps = 1024
            _lib.call("icsg3d_probe_mma_rate", ctypes.c_void_p(out.data_ptr()), 128, n, reps, 2, swz, a_step, b_step, None)
            torch.cuda.synchronize()
            o = out.cpu().tolist()
            d = dict(swz=swz, m=128, n=n, a_step=a_step, b_step=b_step, cyc_per_mma=o[1] / reps,
                     smem_bytes_per_clk=(128 + n) * 32 / (o[1] / reps))
            res.append(d)
            print(f"swz={swz:3d} n={n:3d} a_step={a_step:5d} b_step={b_step:5d} cyc/mma={d['cyc_per_mma']:.1f} smemB/clk={d['smem_bytes_per_clk']:.0f}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mma_rate_probe2.json", "w"), indent=1)
