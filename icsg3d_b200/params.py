"""Parameter containers: Keras-named tensors stored as views of flat fp32 device buffers.

The flat layout is what makes the optimiser one fused kernel launch and the data-parallel gradient
exchange ONE all-reduce (3.36 MB for the VAE).  Names/shapes follow the Keras layers of
vae/lattice_vae.py:160-230 and unet/unet.py:272-355; kernels keep the Keras layouts
(Conv3D (kd,kh,kw,Cin,Cout), Dense (in,out)) so weights can be exchanged with the reference 1:1.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

VAE_FILTERS = [16, 32, 64, 128]   # lattice_vae.py:93
LATENT = 256                      # lattice_vae.py:94
UNET_BLOCKS = [
    ("c1", None, 32), ("c2", 32, 64), ("c3", 64, 64), ("c4", 64, 128), ("c5", 128, 128), ("c6", 128, 256),
    ("c9", 256, 512), ("c10", 512, 512), ("c13", 768, 512), ("c14", 512, 256), ("c15", 384, 256),
    ("c16", 256, 128), ("c17", 192, 128), ("c18", 128, 128),
]


def _conv(specs, name, cin, cout, k=3):
    specs.append((name + "/kernel", (k, k, k, cin, cout), True, ("glorot", k ** 3 * cin, k ** 3 * cout)))
    specs.append((name + "/bias", (cout,), True, ("zeros",)))


def _bn(specs, name, c):
    specs.append((name + "/gamma", (c,), True, ("ones",)))
    specs.append((name + "/beta", (c,), True, ("zeros",)))
    specs.append((name + "/moving_mean", (c,), False, ("zeros",)))
    specs.append((name + "/moving_variance", (c,), False, ("ones",)))


def _dense(specs, name, cin, cout):
    specs.append((name + "/kernel", (cin, cout), True, ("glorot", cin, cout)))
    specs.append((name + "/bias", (cout,), True, ("zeros",)))


def vae_specs(channels=4, ncond=10, d=32, latent=LATENT, filters=VAE_FILTERS):
    """Layer list of LatticeDFCVAE.build_encoder/build_decoder (lattice_vae.py:160-230)."""
    s = []
    cin = channels + channels * ncond  # K.tile repeats the one-hot input_shape[-1] times
    for i, f in enumerate(filters, 1):
        _conv(s, f"enc_conv{i}", cin, f)
        _bn(s, f"enc_bn{i}", f)
        cin = f
    _conv(s, "enc_conv5", cin, 4)
    e = d // 16
    _dense(s, "enc_dense", e * e * e * 4, latent)
    _dense(s, "z_mean", latent, latent)
    _dense(s, "z_log_var", latent, latent)
    e0 = d // 8
    _dense(s, "dec_dense", latent + ncond, e0 * e0 * e0 * 4)
    cin = 4
    for i, f in enumerate(filters[::-1], 1):
        _conv(s, f"dec_conv{i}", cin, f)
        _bn(s, f"dec_bn{i}", f)
        cin = f
    _conv(s, "decoder_output", cin, channels)
    _bn(s, "dec_bn5", channels)
    return s


def unet_specs(channels=4, classes=95):
    """Layer list of AtomUnet.unet_3d_multiclass (unet.py:272-355)."""
    s = []
    for name, cin, cout in UNET_BLOCKS:
        _conv(s, name, channels if cin is None else cin, cout)
        _bn(s, "bn_" + name, cout)
    _conv(s, "soft", 128, classes, k=1)
    _conv(s, "sig", 128, 1, k=1)
    return s


class ParamStore:
    """Named fp32 parameters as views into two flat device buffers (trainable / non-trainable)."""

    PAD = 64  # slack so 16-wide vector reads of a short bias never leave the allocation

    def __init__(self, specs, device, with_grads=True, with_adam=True):
        self.specs = specs
        self.device = torch.device(device)
        nt = sum(int(np.prod(sh)) for _, sh, tr, _ in specs if tr)
        ns = sum(int(np.prod(sh)) for _, sh, tr, _ in specs if not tr)
        self.n_trainable = nt
        self.theta = torch.zeros(nt + self.PAD, dtype=torch.float32, device=self.device)
        self.state = torch.zeros(ns + self.PAD, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(nt + self.PAD, dtype=torch.float32, device=self.device) if with_grads else None
        self.adam_m = torch.zeros(nt + self.PAD, dtype=torch.float32, device=self.device) if with_adam else None
        self.adam_v = torch.zeros(nt + self.PAD, dtype=torch.float32, device=self.device) if with_adam else None
        self.adam_state = torch.zeros(2, dtype=torch.float64, device=self.device) if with_adam else None
        self.p, self.g = OrderedDict(), OrderedDict()
        ot = os_ = 0
        for name, shape, tr, _ in specs:
            n = int(np.prod(shape))
            if tr:
                self.p[name] = self.theta[ot:ot + n].view(shape)
                if with_grads:
                    self.g[name] = self.grad[ot:ot + n].view(shape)
                ot += n
            else:
                self.p[name] = self.state[os_:os_ + n].view(shape)
                os_ += n

    def names(self, trainable=None):
        return [n for n, _, tr, _ in self.specs if trainable is None or tr == trainable]

    def init(self, seed=1):
        """Keras initialisers: glorot_uniform kernels, zero biases, BN gamma=1/beta=0/mean=0/var=1 (SURVEY R3,R5,R8)."""
        gen = torch.Generator().manual_seed(seed)
        for name, shape, _, how in self.specs:
            if how[0] == "glorot":
                limit = math.sqrt(6.0 / (how[1] + how[2]))
                t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit
            elif how[0] == "ones":
                t = torch.ones(shape, dtype=torch.float64)
            else:
                t = torch.zeros(shape, dtype=torch.float64)
            self.p[name].copy_(t.to(torch.float32))
        return self

    def load_dict(self, d, strict=True):
        for name in self.p:
            if name in d:
                t = torch.as_tensor(np.asarray(d[name]) if not torch.is_tensor(d[name]) else d[name])
                if tuple(t.shape) != tuple(self.p[name].shape):
                    raise ValueError(f"{name}: shape {tuple(t.shape)} != {tuple(self.p[name].shape)}")
                self.p[name].copy_(t.to(torch.float32))
            elif strict:
                raise KeyError(f"missing parameter {name}")
        return self

    def to_dict(self):
        return OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in self.p.items())

    def count_trainable(self):
        return self.n_trainable
