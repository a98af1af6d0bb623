"""tcgen05.mma cycles for MN-major operands (the filter-gradient kernels' form) vs K-major, M=128, K=16, bf16."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icsg3d_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
res = []
for n in (64, 128, 256):
    for mn in (0, 1):
        for nacc in (1, 2):
            if nacc * n > 512:
                continue
            _lib.call("icsg3d_probe_mma_rate_mn", ctypes.c_void_p(out.data_ptr()), n, 2048, nacc, mn, None)
            torch.cuda.synchronize()
            o = out.cpu().tolist()
            res.append(dict(n=n, mn_major=mn, nacc=nacc, cyc_per_mma=o[1] / o[0]))
            print(f"n={n:3d} mn_major={mn} nacc={nacc} cycles/mma={o[1] / o[0]:.1f}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mma_rate_probe_mn.json", "w"), indent=1)
