"""Per-entry-point device time of one eager VAE+DFC train step (B=32, 32^3), warm, in the real launch order
(no side-stream overlap so that each call is timed alone).  usage: profile_step.py [batch] [reps] [tag] [d]"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib, utils
from icsg3d_b200.engine import VAEEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tag = sys.argv[3] if len(sys.argv) > 3 else "step"
d = int(sys.argv[4]) if len(sys.argv) > 4 else 32
eng = VAEEngine(B, d=d, seed=1)
eng.overlap_pm = False
eng.overlap_wgrad = False
M, cond, _ = utils.synthetic_batch(B, d=d, seed=1000)
eng.set_inputs(M, cond, torch.randn(B, 256, device="cuda"))
for _ in range(3):
    eng._train_body()
torch.cuda.synchronize()
_lib.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    torch.cuda._sleep(40_000_000)  # queue the whole (CPU-launch-bound) eager step behind a spin: events bracket kernels only
    eng._train_body()
e1.record()
torch.cuda.synchronize()
rec, _lib.PROFILE = _lib.PROFILE, None
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for name, a, b in rec:
    t = a.elapsed_time(b) * 1e3
    agg[name][0] += 1
    agg[name][1] += t
    seq.append((name, t))
tot = sum(v[1] for v in agg.values()) / reps
out = {"batch": B, "eager_ms_per_step": e0.elapsed_time(e1) / reps, "sum_kernel_us_per_step": tot,
       "by_entry": {k: {"n": v[0] // reps, "us": v[1] / reps} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
       "sequence_last_rep": seq[-len(seq) // reps:]}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/profile_{tag}.json", "w"), indent=1)
print(f"eager {out['eager_ms_per_step']:.3f} ms/step, sum of calls {tot:.1f} us")
for k, v in out["by_entry"].items():
    print(f"{k:36s} n={v['n']:3d} {v['us']:9.1f} us {100 * v['us'] / tot:5.1f}%")
