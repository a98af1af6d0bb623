"""Per-entry-point device time of one eager generate.py batch (decode + lattice params + U-Net + argmax/threshold),
warm, each call timed alone.  usage: profile_inference.py [batch] [reps] [tag]"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import _lib
from icsg3d_b200.pipeline import GeneratePipeline
from icsg3d_b200.unet.unet import AtomUnet
from icsg3d_b200.vae.lattice_vae import LatticeDFCVAE

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tag = sys.argv[3] if len(sys.argv) > 3 else "inference"
dev = torch.device("cuda:0")
vae = LatticeDFCVAE(perceptual_model=None, device=dev, seed=1)
vae._set_model(batch_size=B)
unet = AtomUnet(device=dev, seed=2)
pipe = GeneratePipeline(vae, unet, B, use_cuda_graph=False)
z = torch.randn(B, 256, device=dev) * 0.5
cond = torch.eye(10, device=dev)[torch.randint(0, 10, (B,), device=dev)]
for _ in range(3):
    pipe.run(z, cond)
torch.cuda.synchronize()
_lib.PROFILE = []
for _ in range(reps):
    torch.cuda._sleep(40_000_000)
    pipe.run(z, cond)
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
out = {"batch": B, "sum_kernel_us_per_step": tot,
       "by_entry": {k: {"n": v[0] // reps, "us": v[1] / reps} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
       "sequence_last_rep": seq[-len(seq) // reps:]}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/profile_{tag}.json", "w"), indent=1)
print(f"sum of calls {tot:.1f} us per batch of {B}")
for k, v in out["by_entry"].items():
    print(f"{k:36s} n={v['n']:3d} {v['us']:9.1f} us {100 * v['us'] / tot:5.1f}%")
print("sequence:")
for n, t in out["sequence_last_rep"]:
    print(f"  {n:36s} {t:9.1f}")
