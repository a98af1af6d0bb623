"""Device-resident form of the reference's generation loop (generate.py:202-225):

    z ~ N(z_mu_base, var)  ->  decoder.predict([z, cond])  ->  to_lattice_params / to_voxel_params on the
    coordinate channels  ->  unet.model.predict  ->  argmax species labels, sigmoid >= 0.8 atom mask

The reference moves every intermediate through the host (float32 (n,32,32,32,4) decoder output down, back up for the
U-Net, 95-channel probabilities down).  Here the decoder output never leaves the GPU: one fused pass reads it once,
emits the per-sample min/max of the coordinate channels AND the packed bf16 U-Net input (csrc/post.cu), the segmentation
runs on that, and only the results the CPU tail of generate.py needs come back — species labels and mask (uint8),
density channel (fp32), lattice / voxel parameters.  The whole step is one CUDA graph.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops


class GeneratePipeline:
    def __init__(self, vae, unet, batch, threshold=0.8, eps_frac=0.25, use_cuda_graph=True):
        self.B, self.threshold, self.eps_frac = batch, float(threshold), float(eps_frac)
        self.veng = vae.engine(batch)
        self.ueng = unet.engine(batch, train=False)
        if self.veng.d != self.ueng.d:
            raise ValueError("decoder and U-Net grids differ")
        dev = self.veng.dev
        vox = self.veng.d ** 3
        self.nsplit = int(_lib.lib().icsg3d_lattice_nsplit(batch, ctypes.c_int64(vox)))
        self.part = torch.empty(batch, self.nsplit, 6, dtype=torch.float32, device=dev)
        self.lp = torch.empty(batch, 3, dtype=torch.float32, device=dev)
        self.dv = torch.empty(batch, 3, dtype=torch.float32, device=dev)
        self.use_graph, self._graph = use_cuda_graph, None
        self.refresh_weights()

    def refresh_weights(self):
        """Re-pack the GEMM operand copies of both networks (after loading / changing weights)."""
        self.veng.pack_weights()
        self.ueng.pack_weights(dgrad=False)
        self.ueng.inference_coeffs()
        self._graph = None

    def _body(self):
        v, u = self.veng, self.ueng
        v.decode(False)                                                  # generate.py:208 (learning phase 0)
        p = ops._ptr
        vox = v.d ** 3
        _lib.call("icsg3d_coord_minmax", p(v.xhat), 1, 4, 1, self.B, ctypes.c_int64(vox), self.nsplit, p(self.part),
                  p(u.x16), ops._stream())                               # generate.py:211-214 + U-Net input pack
        _lib.call("icsg3d_lattice_finalize", p(self.part), 1, self.B, self.nsplit, ctypes.c_double(self.eps_frac), v.d,
                  p(self.lp), p(self.dv), ops._stream())                 # generate.py:214-217
        u.predict(x_packed=True, threshold=self.threshold, repack=False)  # generate.py:220-225

    def run(self, z, cond):
        """z (B,latent), cond (B,ncond): host or device arrays.  Returns device tensors (views of static buffers, valid
        until the next run): density (B,d,d,d) fp32, species (B,d,d,d) uint8, mask (B,d,d,d) uint8, lattice (B,3),
        voxel (B,3)."""
        v, u = self.veng, self.ueng
        v.z.copy_(torch.as_tensor(z).to(torch.float32), non_blocking=True)
        v.cond.copy_(torch.as_tensor(cond).to(torch.float32), non_blocking=True)
        if not self.use_graph:
            self._body()
        else:
            if self._graph is None:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._body()
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._body()
                self._graph = g
            self._graph.replay()
        return {"density": v.xhat[..., 0], "species": u.argmax, "mask": u.mask, "lattice": self.lp, "voxel": self.dv,
                "M_prime": v.xhat}
