import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icsg3d_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
res = []
for swz in (128, 64):
    for m in (128, 64):
        for n in (16, 64, 128, 256):
            for (a_step, b_step) in ((0, 0), (4096, 0), (0, 2048), (4096, 2048), (swz * 3, 0), (4096 + swz, 2048)):
                reps = 1024
                _lib.call("icsg3d_probe_mma_rate", ctypes.c_void_p(out.data_ptr()), m, n, reps, 2, swz, a_step, b_step, None)
                torch.cuda.synchronize()
                o = out.cpu().tolist()
                res.append(dict(swz=swz, m=m, n=n, a_step=a_step, b_step=b_step, cyc_per_mma=o[1]/reps, mac_per_clk=m*n*16/(o[1]/reps)))
                d = res[-1]
                print(f"swz={swz:3d} m={m:3d} n={n:3d} a_step={a_step:5d} b_step={b_step:5d} cyc/mma={d['cyc_per_mma']:.1f} mac/clk={d['mac_per_clk']:.0f}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mma_rate_probe.json", "w"), indent=1)
