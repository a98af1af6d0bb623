"""Forward executors in the "fp32-class" operand mode (north_star: per-layer activations within 1e-4, losses within
1e-3, U-Net argmax labels bit-exact).

Every conv operand value is carried as a bf16 pair (hi, lo = v - hi); activations are stored [hi | lo | hi] with 3x the
channels and weights [w_hi | w_hi | w_lo] along Cin (csrc/split3.cu), so the ORDINARY tcgen05 conv kernels with fp32
output compute x_hi*w_hi + x_lo*w_hi + x_hi*w_lo — no new conv kernel, 3x the MMA work.  Conv outputs, BatchNorm
statistics, Dense layers and losses are fp32/fp64 as in the bf16 engine.  This is the parity mode: forward only
(activations, losses, labels); it allocates per call and is not the throughput path.

    VAEForwardX3   — encoder, sampling, decoder, both perceptual prefixes and the 4 losses (lattice_vae.py:160-270)
    UNetForwardX3  — learning-phase-0 segmentation pass of generate.py:220-225 (unet.py:272-355)
"""
from __future__ import annotations

import torch

from . import ops
from .engine import PM_BLOCKS
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, POST_NONE, POST_POOL2, POST_UP2, pad16
from .params import LATENT, VAE_FILTERS, ParamStore
from .unet_engine import CAT_CH, UNET_PLAN

BF16, F32, F64 = torch.bfloat16, torch.float32, torch.float64


def unsplit(y3, c, ctot=None, coff=0):
    """fp32 value of a split tensor's channels [coff, coff+c): hi + lo (the buffer holds raw 2-byte values)."""
    ctot = ctot or y3.shape[-1] // 3
    v = y3.view(torch.float16) if ops.SPLIT_FMT == 1 else y3
    return v[..., coff:coff + c].float() + v[..., ctot + coff:ctot + coff + c].float()


class _X3Base:
    def __init__(self, dev):
        self.dev = torch.device(dev)
        self._part = torch.zeros(148 * 4 * 2 * 1024, dtype=F64, device=self.dev)

    def conv(self, x3, w, bias, cout, act=ACT_NONE, cin_pad=None, fold=None):
        """3x3x3 (or 1x1x1) conv of a split tensor -> fp32 [B,D,H,W,cout]."""
        kw = dict(cin_lead=fold[0], fold=fold[1], fold_c=fold[2]) if fold else {}
        wp = ops.pack_conv_w_fprop_x3(w, cin_pad=cin_pad, **kw)
        B, D, H, W, _ = x3.shape
        out = torch.empty(B, D, H, W, cout, dtype=F32, device=self.dev)
        ops.conv3d_k3(x3, wp, bias, out=out, act=act, n_store=cout, split=True)
        return out

    def bn_coeffs(self, x32, C, gamma, beta, mm, mv, training):
        scale, shift = torch.empty(C, device=self.dev), torch.empty(C, device=self.dev)
        if training:  # batch statistics (the moving averages are not touched by this parity pass)
            rows = x32.numel() // C
            n = ops.bn_nparts(rows, C, x32.dtype)
            part = self._part[: n * 2 * C].view(n, 2, C)
            ops.bn_stats(x32, C, part)
            sums = torch.empty(2 * C, dtype=F64, device=self.dev)
            mean, rstd = torch.empty(C, device=self.dev), torch.empty(C, device=self.dev)
            ops.bn_reduce_finalize(part, float(rows), gamma, beta, sums, mean, rstd, scale, shift)
        else:
            ops.bn_inference_coeffs(gamma, beta, mm, mv, scale, shift)
        return scale, shift

    def bn_split(self, x32, C, scale, shift, act, post, y3=None, ctot=None, coff=0):
        """BatchNorm apply (+activation, +pool / upsample) of an fp32 tensor into a split tensor."""
        B, D, H, W, _ = x32.shape
        ctot = ctot or pad16(C)
        if y3 is None:
            Do = D // 2 if post == POST_POOL2 else (2 * D if post == POST_UP2 else D)
            y3 = torch.zeros(B, Do, Do, Do, 3 * ctot, dtype=BF16, device=self.dev)
        ops.bn_apply_fwd_split3(x32, C, scale, shift, act, post, y3, ctot, coff)
        return y3


class VAEForwardX3(_X3Base):
    def __init__(self, batch, d=32, ncond=10, latent=LATENT, filters=VAE_FILTERS, device="cuda",
                 vae_params: ParamStore = None, pm_params: ParamStore = None, alpha=0.5, beta=3e-4,
                 pm_layer_weights=(1.0, 1.0, 1.0, 1.0)):
        super().__init__(device)
        self.B, self.d, self.ncond, self.latent, self.filters = batch, d, ncond, latent, list(filters)
        self.vp, self.pp = vae_params, pm_params
        self.alpha, self.beta, self.pm_w = float(alpha), float(beta), [float(w) for w in pm_layer_weights]
        self.taps = {}

    def pm_forward(self, x3, training, prefix):
        """U-Net prefix c1..c10 on a split input; returns the 4 DFC tap tensors (post-ReLU, fp32)."""
        p = self.pp.p
        feats = []
        for name, cin, cout, lvl, pool, tap in PM_BLOCKS:
            a = self.conv(x3, p[name + "/kernel"], p[name + "/bias"], cout, act=ACT_RELU, cin_pad=pad16(cin))
            self.taps[prefix + name] = a
            if tap:
                feats.append(a)
            if name == "c10":
                break
            bn = "bn_" + name
            sc, sh = self.bn_coeffs(a, cout, p[bn + "/gamma"], p[bn + "/beta"], p[bn + "/moving_mean"],
                                    p[bn + "/moving_variance"], training)
            x3 = self.bn_split(a, cout, sc, sh, ACT_NONE, POST_POOL2 if pool else POST_NONE)
        return feats

    def forward(self, M, cond, eps, training=True):
        """-> [loss, pm, mse, kld] (device fp32 tensor); per-layer fp32 activations in self.taps (oracle tap names)."""
        p, B, d, dev = self.vp.p, self.B, self.d, self.dev
        M, cond, eps = M.to(dev, F32).contiguous(), cond.to(dev, F32).contiguous(), eps.to(dev, F32).contiguous()
        t = self.taps = {}
        xe3 = torch.zeros(B, d, d, d, 48, dtype=BF16, device=dev)
        xp3 = torch.zeros(B, d, d, d, 48, dtype=BF16, device=dev)
        ops.pack_vae_input_split3(M, cond, xe3, xp3)
        # ---- encoder (lattice_vae.py:160-195) ----
        x3 = xe3
        for i, f in enumerate(self.filters, 1):
            name, bn = f"enc_conv{i}", f"enc_bn{i}"
            c = self.conv(x3, p[name + "/kernel"], p[name + "/bias"], f, cin_pad=16 if i == 1 else None,
                          fold=(4, 4, self.ncond) if i == 1 else None)
            t[name] = c
            sc, sh = self.bn_coeffs(c, f, p[bn + "/gamma"], p[bn + "/beta"], p[bn + "/moving_mean"], p[bn + "/moving_variance"],
                                    training)
            x3 = self.bn_split(c, f, sc, sh, ACT_LEAKY, POST_POOL2)
            t[f"enc_pool{i}"] = unsplit(x3, f)
        e5 = self.conv(x3, p["enc_conv5/kernel"], p["enc_conv5/bias"], 4, act=ACT_LEAKY)
        t["enc_conv5"] = e5
        z = lambda *s: torch.empty(*s, dtype=F32, device=dev)
        h, mu, lv, zz, kl = z(B, self.latent), z(B, self.latent), z(B, self.latent), z(B, self.latent), z(B)
        ops.dense_fwd(e5.view(B, -1), p["enc_dense/kernel"], p["enc_dense/bias"], h, act=ACT_RELU)
        ops.dense_fwd(h, p["z_mean/kernel"], p["z_mean/bias"], mu)
        ops.dense_fwd(h, p["z_log_var/kernel"], p["z_log_var/bias"], lv)
        ops.reparam_fwd(mu, lv, eps, zz, kl)
        t.update(z_mean=mu, z_log_var=lv, z=zz)
        # ---- decoder (lattice_vae.py:197-230) ----
        s0 = d // 8
        dd = z(B, s0 * s0 * s0 * 4)
        ops.dense_fwd(zz, p["dec_dense/kernel"], p["dec_dense/bias"], dd, x2=cond)
        x3 = torch.zeros(B, s0, s0, s0, 48, dtype=BF16, device=dev)
        ops.f32_to_split3(dd.view(B, s0, s0, s0, 4), 4, x3, 16)
        nf = len(self.filters)
        for i, f in enumerate(self.filters[::-1], 1):
            name, bn = f"dec_conv{i}", f"dec_bn{i}"
            c = self.conv(x3, p[name + "/kernel"], p[name + "/bias"], f, cin_pad=16 if i == 1 else None)
            t[name] = c
            sc, sh = self.bn_coeffs(c, f, p[bn + "/gamma"], p[bn + "/beta"], p[bn + "/moving_mean"], p[bn + "/moving_variance"],
                                    training)
            x3 = self.bn_split(c, f, sc, sh, ACT_LEAKY, POST_UP2 if i < nf else POST_NONE)
        c5 = self.conv(x3, p["decoder_output/kernel"], p["decoder_output/bias"], 4)
        t["decoder_output"] = c5
        sc, sh = self.bn_coeffs(c5, 4, p["dec_bn5/gamma"], p["dec_bn5/beta"], p["dec_bn5/moving_mean"],
                                p["dec_bn5/moving_variance"], training)
        xhat = z(B, d, d, d, 4)
        ops.bn_apply_fwd(c5, 4, sc, sh, ACT_RELU, POST_NONE, y32=xhat)
        t["x_hat"] = xhat
        xh3 = torch.zeros(B, d, d, d, 48, dtype=BF16, device=dev)
        ops.f32_to_split3(xhat, 4, xh3, 16)
        # ---- perceptual prefixes + losses (lattice_vae.py:232-270) ----
        fx = self.pm_forward(xp3, training, "pm_x/")
        fh = self.pm_forward(xh3, training, "pm_xhat/")
        terms = [(M, xhat, 1.0 / M.numel())] + [(a, b, w / a.numel()) for a, b, w in zip(fx, fh, self.pm_w)]
        stride = 148 * 4
        partials = torch.zeros(len(terms), stride, dtype=F64, device=dev)
        nparts = [ops.sqdiff_nparts(a.numel()) for a, _, _ in terms]
        for k, (a, b, _) in enumerate(terms):
            ops.sqdiff_partials(a, b, partials[k], nparts[k])
        metrics = torch.zeros(4, dtype=F32, device=dev)
        ops.vae_loss_assemble(partials, torch.tensor(nparts, dtype=torch.int32, device=dev),
                              torch.tensor([s for _, _, s in terms], dtype=F64, device=dev), kl, 1.0 / B, self.alpha, self.beta,
                              metrics)
        return metrics


class UNetForwardX3(_X3Base):
    def __init__(self, batch, d=32, channels=4, classes=95, device="cuda", params: ParamStore = None):
        super().__init__(device)
        assert channels == 4
        self.B, self.d, self.classes, self.pp = batch, d, classes, params
        self.taps = {}

    def predict(self, X):
        """-> (logits fp32 [B,d,d,d,96] with the sigmoid logit in column `classes`, argmax uint8, sigmoid prob fp32)."""
        p, B, d, dev = self.pp.p, self.B, self.d, self.dev
        X = X.to(dev, F32).contiguous()
        x3 = torch.zeros(B, d, d, d, 48, dtype=BF16, device=dev)
        ops.pack_vae_input_split3(X, None, None, x3)
        cat = {k: torch.zeros(B, d >> lvl, d >> lvl, d >> lvl, 3 * c, dtype=BF16, device=dev) for k, (lvl, c) in CAT_CH.items()}
        y3, p3 = {}, {}
        for L in UNET_PLAN:
            n, cout, src = L["n"], L["cout"], L["src"]
            xin = (x3 if src == "x" else p3[src[5:]] if src.startswith("pool:") else cat[src[4:]] if src.startswith("cat:")
                   else y3[src])
            cin = 4 if n == "c1" else L["cin"]
            a = self.conv(xin, p[n + "/kernel"], p[n + "/bias"], cout, act=ACT_RELU, cin_pad=pad16(cin))
            self.taps[n] = a
            sc, sh = self.bn_coeffs(a, cout, p[f"bn_{n}/gamma"], p[f"bn_{n}/beta"], p[f"bn_{n}/moving_mean"],
                                    p[f"bn_{n}/moving_variance"], False)
            if "up" in L:
                buf, off = L["up"]
                self.bn_split(a, cout, sc, sh, ACT_NONE, POST_UP2, y3=cat[buf], ctot=CAT_CH[buf][1], coff=off)
            elif "cat" in L:
                buf, off = L["cat"]
                self.bn_split(a, cout, sc, sh, ACT_NONE, POST_NONE, y3=cat[buf], ctot=CAT_CH[buf][1], coff=off)
            else:
                y3[n] = self.bn_split(a, cout, sc, sh, ACT_NONE, POST_NONE)
            if L.get("pool"):
                p3[n] = self.bn_split(a, cout, sc, sh, ACT_NONE, POST_POOL2)
        nh = pad16(self.classes + 1)
        wh = torch.cat([p["soft/kernel"], p["sig/kernel"]], dim=-1).contiguous()  # (1,1,1,128,classes+1)
        bh = torch.zeros(nh, dtype=F32, device=dev)
        bh[: self.classes] = p["soft/bias"]
        bh[self.classes] = p["sig/bias"][0]
        logits = self.conv(y3["c18"], wh, bh, nh)
        M = B * d ** 3
        argmax = torch.empty(B, d, d, d, dtype=torch.uint8, device=dev)
        sigp = torch.empty(B, d, d, d, dtype=F32, device=dev)
        species = torch.zeros(B, d, d, d, dtype=torch.uint8, device=dev)
        partials = torch.zeros(ops.heads_loss_nparts(M), 6, dtype=F64, device=dev)
        class_w = torch.full((self.classes,), float(self.classes), device=dev)
        ops.heads_loss(logits, self.classes, species, class_w, 1.0 / M, partials, argmax_out=argmax, sig_prob=sigp)
        return logits, argmax, sigp
