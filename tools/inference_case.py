"""Run the generate.py batch a few times eagerly (for ncu).  usage: inference_case.py [batch] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icsg3d_b200.pipeline import GeneratePipeline
from icsg3d_b200.unet.unet import AtomUnet
from icsg3d_b200.vae.lattice_vae import LatticeDFCVAE
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
vae = LatticeDFCVAE(perceptual_model=None, device=dev, seed=1)
vae._set_model(batch_size=B)
pipe = GeneratePipeline(vae, AtomUnet(device=dev, seed=2), B, use_cuda_graph=False)
z = torch.randn(B, 256, device=dev) * 0.5
cond = torch.eye(10, device=dev)[torch.randint(0, 10, (B,), device=dev)]
for _ in range(iters):
    pipe.run(z, cond)
torch.cuda.synchronize()
