"""Secondary measurements (not the contract line): configs[0] U-Net train step (B=8, 32^3), configs[4] inference
(decoder + U-Net + argmax/threshold) and voxeliser throughput, configs[3] 64^3.  Prints one JSON object."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from icsg3d_b200 import ops, utils
from icsg3d_b200.engine import VAEEngine
from icsg3d_b200.unet_engine import UNetEngine


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
# ---- voxeliser ----
n = 512
sites, nsites, lat = utils.synthetic_cells(n, seed=1)
ms = timeit(lambda: utils.voxelize_cells(sites, nsites, lat, d=32), n=10)
out["voxeliser_32"] = {"samples_per_s": n / ms * 1e3, "gbps": n * 557056 / ms / 1e6, "ms_per_512": ms}
# ---- U-Net train step, B=8, 32^3 (configs[0]) ----
M, cond, S = utils.synthetic_batch(8, d=32, seed=2)
un = UNetEngine(8, d=32, lr=3e-6)
un.set_inputs(M, S)
un.capture_train_graph()
ms = timeit(un.train_step, n=10)
out["unet_train_B8_32"] = {"ms_per_step": ms, "samples_per_s": 8 / ms * 1e3, "conv_tflops": 377.9 * 8 / ms}
del un
torch.cuda.empty_cache()
# ---- inference: decoder + U-Net + labels (configs[4]) ----
B = 64
vae = VAEEngine(B, d=32)
uni = UNetEngine(B, d=32, train=False)
vae.z.normal_()
vae.cond.zero_()
vae.cond[:, 0] = 1


def infer():
    vae.pack_weights()
    vae.decode(False)
    uni.X.copy_(vae.xhat)
    uni.predict()


ms = timeit(infer, n=5)
out["inference_decode_segment_32"] = {"ms_per_batch64": ms, "samples_per_s": B / ms * 1e3, "tflops": 127.76 * B / ms}
del vae, uni
torch.cuda.empty_cache()
# ---- 64^3 (configs[3]): U-Net train B=16 would need ~40 GB of activations: run B=4, report per-sample ----
try:
    M, cond, S = utils.synthetic_batch(4, d=64, seed=3)
    un = UNetEngine(4, d=64, lr=3e-6)
    un.set_inputs(M, S)
    ms = timeit(un.train_step, n=3, warm=2)
    out["unet_train_B4_64"] = {"ms_per_step": ms, "samples_per_s": 4 / ms * 1e3, "conv_tflops": 3023.5 * 4 / ms}
    del un
    torch.cuda.empty_cache()
    ve = VAEEngine(16, d=64)
    M, cond, _ = utils.synthetic_batch(16, d=64, seed=4)
    ve.set_inputs(M, cond)
    ms = timeit(ve.train_step, n=3, warm=2)
    out["vae_dfc_train_B16_64_extension"] = {"ms_per_step": ms, "samples_per_s": 16 / ms * 1e3, "conv_tflops": 288.4 * 16 / ms}
except Exception as e:  # noqa: BLE001
    out["64cubed_error"] = repr(e)[:300]
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_extra.json", "w"), indent=1)
