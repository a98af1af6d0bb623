"""Per-entry-point device time of one eager U-Net train step (configs[0]: B=8, 32^3), warm, each call timed alone.
usage: profile_unet_train.py [batch] [d] [reps]"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib, utils
from icsg3d_b200.unet_engine import UNetEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
eng = UNetEngine(B, d=d, device=dev, seed=2)
eng.overlap_wgrad = False  # each call timed alone
M, _, S = utils.synthetic_batch(B, d=d, seed=2000, device=dev)
eng.set_inputs(M, S)
for _ in range(3):
    eng._train_body()
torch.cuda.synchronize()
_lib.PROFILE = []
for _ in range(reps):
    torch.cuda._sleep(40_000_000)
    eng._train_body()
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
print(f"sum of calls {tot:.1f} us per step (B={B}, d={d})")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:36s} n={v[0] // reps:3d} {v[1] / reps:9.1f} us {100 * v[1] / reps / tot:5.1f}%")
print("sequence:")
for n, t in seq[-len(seq) // reps:]:
    print(f"  {n:36s} {t:9.1f}")
