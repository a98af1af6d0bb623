import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icsg3d_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
res = []
# (name, G, nt, plane_rows, WP, row_bytes, ksteps)
cases = [("c2f", 3, 64, 330, 33, 64, 2), ("c2d", 3, 32, 330, 33, 128, 4), ("c1f", 5, 32, 594, 33, 32, 1),
         ("enc1", 14, 16, 594, 33, 32, 1)]
for name, G, nt, pr, WP, rb, ks in cases:
    for mode in (0, 4):
        _lib.call("icsg3d_probe_halo_pattern", ctypes.c_void_p(out.data_ptr()), G, nt, pr, WP, rb, ks, 20, mode, None)
        torch.cuda.synchronize()
        o = out.cpu().tolist()
        r = dict(case=name, G=G, nt=nt, row_bytes=rb, ksteps=ks, mode=mode, cyc_per_mma=o[0] / o[1])
        res.append(r)
        print(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/halo_pattern_probe.json", "w"), indent=1)
