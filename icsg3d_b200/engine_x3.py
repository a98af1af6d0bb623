"""Executors in the "fp32-class" operand mode (north_star: per-layer activations AND gradients within 1e-4, losses
within 1e-3, U-Net argmax labels bit-exact) — the parity mode behind `LatticeDFCVAE(dtype="fp32")` /
`AtomUnet(dtype="fp32")`.

Every conv operand value is carried as a pair (hi, lo = v - hi) of 2-byte floats; activations are stored
[hi | lo | hi] with 3x the channels and weights [w_hi | w_hi | w_lo] along K (csrc/split3.cu), so the ORDINARY tcgen05
conv kernels with fp32 output compute x_hi*w_hi + x_lo*w_hi + x_hi*w_lo — no new conv kernel, 3x the MMA work.
The backward pass uses the same trick on the same kernels:

    dgrad   the conv kernel on dy stored [hi | lo | hi] with the mirrored-tap weight pack [w_hi | w_hi | w_lo] along Cout;
    wgrad   the filter-gradient kernel on the channel slices x = [x_hi | x_lo], dy = [dy_hi | dy_lo]; the 2x2 block
            result is folded hi*hi + lo*hi + hi*lo (csrc/split3.cu::wgrad_combine_x3_kernel);
    BatchNorm backward, loss seeds, Dense layers, Adam: fp32 tensors, fp64 sums.

Inference (`predict`) uses IEEE fp16 pairs (22 significant bits, weights pre-scaled by 2^10); training uses bf16 pairs
(16 significant bits but the full fp32 exponent range — gradients of 1e-8 would be subnormal in fp16).
Conv outputs, BatchNorm statistics, Dense layers and losses are fp32/fp64 as in the bf16 engine.  This mode allocates per
call and is not the throughput path.

    VAEForwardX3 / VAETrainX3   — encoder, sampling, decoder, both perceptual prefixes, the 4 losses, backward, Keras-Adam
                                   (lattice_vae.py:160-270, 296-298)
    UNetForwardX3 / UNetTrainX3 — the segmentation pass of generate.py:220-225 and the U-Net train step (unet.py:272-355)
"""
from __future__ import annotations

import torch

from . import ops
from .engine import PM_BLOCKS
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, POST_NONE, POST_POOL2, POST_UP2, pad16
from .params import LATENT, VAE_FILTERS, ParamStore
from .unet_engine import CAT_CH, UNET_PLAN

BF16, F32, F64 = torch.bfloat16, torch.float32, torch.float64


def unsplit(y3, c, ctot=None, coff=0, fmt=None):
    """fp32 value of a split tensor's channels [coff, coff+c): hi + lo (the buffer holds raw 2-byte values)."""
    ctot = ctot or y3.shape[-1] // 3
    v = y3.view(torch.float16) if (ops.SPLIT_FMT if fmt is None else fmt) == 1 else y3
    return v[..., coff:coff + c].float() + v[..., ctot + coff:ctot + coff + c].float()


class _X3Base:
    def __init__(self, dev, fmt=None):
        self.dev = torch.device(dev)
        self.fmt = ops.SPLIT_FMT if fmt is None else fmt
        self._part = torch.zeros(148 * 4 * 2 * 1024, dtype=F64, device=self.dev)

    # ---- forward pieces ----
    def conv(self, x3, w, bias, cout, act=ACT_NONE, cin_pad=None, fold=None):
        """3x3x3 (or 1x1x1) conv of a split tensor -> fp32 [B,D,H,W,cout]."""
        kw = dict(cin_lead=fold[0], fold=fold[1], fold_c=fold[2]) if fold else {}
        wp = ops.pack_conv_w_fprop_x3(w, cin_pad=cin_pad, fmt=self.fmt, **kw)
        B, D, H, W, _ = x3.shape
        out = torch.empty(B, D, H, W, cout, dtype=F32, device=self.dev)
        ops.conv3d_k3(x3, wp, bias, out=out, act=act, n_store=cout, split=True, fmt=self.fmt)
        return out

    def bn_coeffs(self, x32, C, gamma, beta, mm, mv, training):
        """-> dict(scale, shift, mean, rstd): batch statistics when training (moving averages untouched)."""
        z = lambda: torch.empty(C, device=self.dev)
        st = dict(scale=z(), shift=z(), mean=z(), rstd=z())
        if training:
            rows = x32.numel() // C
            n = ops.bn_nparts(rows, C, x32.dtype)
            part = self._part[: n * 2 * C].view(n, 2, C)
            ops.bn_stats(x32, C, part)
            sums = torch.empty(2 * C, dtype=F64, device=self.dev)
            ops.bn_reduce_finalize(part, float(rows), gamma, beta, sums, st["mean"], st["rstd"], st["scale"], st["shift"])
        else:
            ops.bn_inference_coeffs(gamma, beta, mm, mv, st["scale"], st["shift"])
        return st

    def bn_split(self, x32, C, st, act, post, y3=None, ctot=None, coff=0, want_idx=False):
        """BatchNorm apply (+activation, +pool / upsample) of an fp32 tensor into a split tensor [, pool arg-max index]."""
        B, D, H, W, _ = x32.shape
        ctot = ctot or pad16(C)
        Do = D // 2 if post == POST_POOL2 else (2 * D if post == POST_UP2 else D)
        if y3 is None:
            y3 = torch.zeros(B, Do, Do, Do, 3 * ctot, dtype=BF16, device=self.dev)
        idx = torch.empty(B, Do, Do, Do, C, dtype=torch.uint8, device=self.dev) if want_idx and post == POST_POOL2 else None
        ops.bn_apply_fwd_split3(x32, C, st["scale"], st["shift"], act, post, y3, ctot, coff, pool_idx=idx, fmt=self.fmt)
        return (y3, idx) if want_idx else y3

    def split(self, x32, c, ctot=None):
        """fp32 [..., ld] (first c channels) -> split tensor [..., 3*ctot]."""
        ctot = ctot or pad16(c)
        y3 = torch.zeros(*x32.shape[:-1], 3 * ctot, dtype=BF16, device=self.dev)
        ops.f32_to_split3(x32, c, y3, ctot, fmt=self.fmt)
        return y3

    # ---- backward pieces (bf16 pairs only: fmt 0) ----
    def bn_bwd(self, dy, x, C, st, act, post, idx, pre_relu=False, tap_other=None, tap_coef=0.0, dgamma=None, dbeta=None,
               dy2=None):
        """fp32 BatchNorm(+activation, +pool/upsample) backward: returns dx (gradient w.r.t. the conv output x)."""
        rows = x.numel() // C
        n = ops.bn_bwd_nparts(x, C, post)
        part = self._part[: n * 2 * C].view(n, 2, C)
        ops.bn_bwd_reduce(dy, x, C, st["mean"], st["rstd"], st["scale"], st["shift"], act, post, idx, part, dy2=dy2)
        sums = torch.empty(2 * C, dtype=F64, device=self.dev)
        ops.bn_reduce_grads(part, sums, dgamma, dbeta)
        dx = torch.empty_like(x)
        ops.bn_bwd_apply_f32(dy, x, C, st["mean"], st["rstd"], st["scale"], st["shift"], act, post, idx, sums, float(rows), dx,
                             pre_relu=pre_relu, tap_other=tap_other, tap_coef=tap_coef, dy2=dy2)
        return dx

    def conv_dgrad(self, dc3, w, cin, cin_pad=None):
        """Input gradient of a conv with Keras kernel w from the split output gradient dc3 -> fp32 [B,D,H,W,cin]."""
        cout_pad = dc3.shape[-1] // 3
        cin_pad = cin_pad or pad16(w.shape[3])
        wd3 = ops.pack_conv_w_dgrad_x3(w, cin_pad=cin_pad, cout_pad=cout_pad, fmt=self.fmt)
        B, D, H, W, _ = dc3.shape
        out = torch.empty(B, D, H, W, cin, dtype=F32, device=self.dev)
        ops.conv3d_k3(dc3, wd3, None, out=out, n_store=cin, split=True, fmt=self.fmt)
        return out

    def conv_wgrad(self, x3, dc3, gk, cin, cout, fold=None, k1=False):
        """Filter gradient into gk (Keras layout) from the split input x3 and the split output gradient dc3."""
        cin_pad, cout_pad = x3.shape[-1] // 3, dc3.shape[-1] // 3
        if k1:  # the two 1x1x1 heads as one [cin x (classes+1)] matrix: returned padded, unpacked by unpack_heads_grad
            scratch = torch.empty(1, cin_pad, cout_pad, dtype=F32, device=self.dev)
            return ops.conv3d_wgrad_x3(x3, dc3, cin_pad, cout_pad, scratch, k1=True)
        if cin == cin_pad and cout == cout_pad and fold is None:
            ops.conv3d_wgrad_x3(x3, dc3, cin_pad, cout_pad, gk.view(27, cin, cout))
            return
        scratch = torch.empty(27, cin_pad, cout_pad, dtype=F32, device=self.dev)
        ops.conv3d_wgrad_x3(x3, dc3, cin_pad, cout_pad, scratch)
        if fold is None:
            ops.unpack_conv_dw(scratch, cin, cout, out=gk)
        else:
            ops.unpack_conv_dw(scratch, cin, cout, cin_lead=fold[0], fold=fold[1], fold_c=fold[2], out=gk)

    def col_sums(self, x32, C, out):
        """out[c] = sum over rows of x32[..., c] (bias gradients), fp64 accumulation."""
        rows = x32.numel() // x32.shape[-1]
        n = ops.bn_nparts(rows, C, x32.dtype)
        part = self._part[: n * 2 * C].view(n, 2, C)
        ops.bn_stats(x32, C, part)
        sums = torch.empty(2 * C, dtype=F64, device=self.dev)
        ops.bn_reduce_grads(part, sums, None, out)
        return sums


class VAEForwardX3(_X3Base):
    def __init__(self, batch, d=32, ncond=10, latent=LATENT, filters=VAE_FILTERS, device="cuda",
                 vae_params: ParamStore = None, pm_params: ParamStore = None, alpha=0.5, beta=3e-4,
                 pm_layer_weights=(1.0, 1.0, 1.0, 1.0), fmt=None):
        super().__init__(device, fmt)
        self.B, self.d, self.ncond, self.latent, self.filters = batch, d, ncond, latent, list(filters)
        self.vp, self.pp = vae_params, pm_params
        self.alpha, self.beta, self.pm_w = float(alpha), float(beta), [float(w) for w in pm_layer_weights]
        self.taps, self.saved = {}, {}

    def pm_forward(self, x3, training, prefix):
        """U-Net prefix c1..c10 on a split input; returns the 4 DFC tap tensors (post-ReLU, fp32)."""
        p = self.pp.p
        feats = []
        for name, cin, cout, lvl, pool, tap in PM_BLOCKS:
            a = self.conv(x3, p[name + "/kernel"], p[name + "/bias"], cout, act=ACT_RELU, cin_pad=pad16(cin))
            self.taps[prefix + name] = a
            rec = self.saved[prefix + name] = dict(x3=x3, a=a)
            if tap:
                feats.append(a)
            if name == "c10":
                break
            bn = "bn_" + name
            st = self.bn_coeffs(a, cout, p[bn + "/gamma"], p[bn + "/beta"], p[bn + "/moving_mean"],
                                p[bn + "/moving_variance"], training)
            x3, idx = self.bn_split(a, cout, st, ACT_NONE, POST_POOL2 if pool else POST_NONE, want_idx=True)
            rec.update(st=st, idx=idx)
        return feats

    def forward(self, M, cond, eps, training=True):
        """-> [loss, pm, mse, kld] (device fp32 tensor); per-layer fp32 activations in self.taps (oracle tap names)."""
        p, B, d, dev = self.vp.p, self.B, self.d, self.dev
        M, cond, eps = M.to(dev, F32).contiguous(), cond.to(dev, F32).contiguous(), eps.to(dev, F32).contiguous()
        t, S = self.taps, self.saved
        t.clear()
        S.clear()
        S.update(M=M, cond=cond, eps=eps)
        xe3 = torch.zeros(B, d, d, d, 48, dtype=BF16, device=dev)
        xp3 = torch.zeros(B, d, d, d, 48, dtype=BF16, device=dev)
        ops.pack_vae_input_split3(M, cond, xe3, xp3, fmt=self.fmt)
        # ---- encoder (lattice_vae.py:160-195) ----
        x3 = xe3
        for i, f in enumerate(self.filters, 1):
            name, bn = f"enc_conv{i}", f"enc_bn{i}"
            c = self.conv(x3, p[name + "/kernel"], p[name + "/bias"], f, cin_pad=16 if i == 1 else None,
                          fold=(4, 4, self.ncond) if i == 1 else None)
            t[name] = c
            st = self.bn_coeffs(c, f, p[bn + "/gamma"], p[bn + "/beta"], p[bn + "/moving_mean"], p[bn + "/moving_variance"],
                                training)
            y3, idx = self.bn_split(c, f, st, ACT_LEAKY, POST_POOL2, want_idx=True)
            S[name] = dict(x3=x3, c=c, st=st, idx=idx)
            x3 = y3
            t[f"enc_pool{i}"] = unsplit(x3, f, fmt=self.fmt)
        e5 = self.conv(x3, p["enc_conv5/kernel"], p["enc_conv5/bias"], 4, act=ACT_LEAKY)
        t["enc_conv5"] = e5
        S["enc_conv5"] = dict(x3=x3, e5=e5)
        z = lambda *s: torch.empty(*s, dtype=F32, device=dev)
        h, mu, lv, zz, kl = z(B, self.latent), z(B, self.latent), z(B, self.latent), z(B, self.latent), z(B)
        ops.dense_fwd(e5.view(B, -1), p["enc_dense/kernel"], p["enc_dense/bias"], h, act=ACT_RELU)
        ops.dense_fwd(h, p["z_mean/kernel"], p["z_mean/bias"], mu)
        ops.dense_fwd(h, p["z_log_var/kernel"], p["z_log_var/bias"], lv)
        ops.reparam_fwd(mu, lv, eps, zz, kl)
        t.update(z_mean=mu, z_log_var=lv, z=zz)
        S.update(h=h, mu=mu, lv=lv, z=zz)
        # ---- decoder (lattice_vae.py:197-230) ----
        s0 = d // 8
        dd = z(B, s0 * s0 * s0 * 4)
        ops.dense_fwd(zz, p["dec_dense/kernel"], p["dec_dense/bias"], dd, x2=cond)
        x3 = self.split(dd.view(B, s0, s0, s0, 4), 4, 16)
        nf = len(self.filters)
        for i, f in enumerate(self.filters[::-1], 1):
            name, bn = f"dec_conv{i}", f"dec_bn{i}"
            c = self.conv(x3, p[name + "/kernel"], p[name + "/bias"], f, cin_pad=16 if i == 1 else None)
            t[name] = c
            st = self.bn_coeffs(c, f, p[bn + "/gamma"], p[bn + "/beta"], p[bn + "/moving_mean"], p[bn + "/moving_variance"],
                                training)
            S[name] = dict(x3=x3, c=c, st=st, up=i < nf)
            x3 = self.bn_split(c, f, st, ACT_LEAKY, POST_UP2 if i < nf else POST_NONE)
        c5 = self.conv(x3, p["decoder_output/kernel"], p["decoder_output/bias"], 4)
        t["decoder_output"] = c5
        st5 = self.bn_coeffs(c5, 4, p["dec_bn5/gamma"], p["dec_bn5/beta"], p["dec_bn5/moving_mean"],
                             p["dec_bn5/moving_variance"], training)
        xhat = z(B, d, d, d, 4)
        ops.bn_apply_fwd(c5, 4, st5["scale"], st5["shift"], ACT_RELU, POST_NONE, y32=xhat)
        t["x_hat"] = xhat
        S["decoder_output"] = dict(x3=x3, c=c5, st=st5, xhat=xhat)
        xh3 = self.split(xhat, 4, 16)
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
        self.metrics = metrics
        return metrics


class VAETrainX3(VAEForwardX3):
    """One fp32-class train_on_batch (lattice_vae.py:296-298): forward, backward into vae_params.g, Keras-Adam."""

    def __init__(self, *a, lr=5e-4, **kw):
        kw["fmt"] = 0  # bf16 pairs: full fp32 exponent range for the gradients
        super().__init__(*a, **kw)
        self.lr = float(lr)

    def backward(self):
        S, p, g, pp, B, d = self.saved, self.vp.p, self.vp.g, self.pp.p, self.B, self.d
        M = S["M"]
        # ---- perceptual branch on x_hat: dgrad only (pm is frozen) ----
        names = [b[0] for b in PM_BLOCKS]
        coef, k = {}, 0
        for name, _, _, _, _, tap in PM_BLOCKS:
            if tap:
                feat = S["pm_xhat/" + name]["a"].numel() // B
                coef[name] = self.alpha * self.pm_w[k] * 2.0 / (feat * B)
                k += 1
        a1, a0 = S["pm_xhat/c10"]["a"], S["pm_x/c10"]["a"]
        dc = torch.empty_like(a1)
        ops.tap_grad_relu_f32(a1, a0, coef["c10"], dc)
        dy = self.conv_dgrad(self.split(dc, 512), pp["c10/kernel"], 512)
        for name, cin, cout, lvl, pool, tap in reversed(PM_BLOCKS[:-1]):
            L = S["pm_xhat/" + name]
            dc = self.bn_bwd(dy, L["a"], cout, L["st"], ACT_NONE, POST_POOL2 if pool else POST_NONE, L["idx"], pre_relu=True,
                             tap_other=S["pm_x/" + name]["a"] if tap else None, tap_coef=coef.get(name, 0.0))
            dy = self.conv_dgrad(self.split(dc, cout), pp[name + "/kernel"], cin)
        dpm = dy  # (B,d,d,d,4)
        # ---- decoder ----
        L5 = S["decoder_output"]
        dxhat = torch.empty_like(L5["xhat"])
        ops.xhat_grad_f32(M, L5["xhat"], 2.0 / (M.numel() // B * B), dpm, dxhat)
        dc5 = self.bn_bwd(dxhat, L5["c"], 4, L5["st"], ACT_RELU, POST_NONE, None, dgamma=g["dec_bn5/gamma"],
                          dbeta=g["dec_bn5/beta"])
        dc3 = self.split(dc5, 4, 16)
        f0 = self.filters[0]
        self.conv_wgrad(L5["x3"], dc3, g["decoder_output/kernel"], f0, 4)
        du = self.conv_dgrad(dc3, p["decoder_output/kernel"], f0)
        nf = len(self.filters)
        for i in range(nf, 0, -1):
            name, bn, f = f"dec_conv{i}", f"dec_bn{i}", self.filters[::-1][i - 1]
            L = S[name]
            dc = self.bn_bwd(du, L["c"], f, L["st"], ACT_LEAKY, POST_UP2 if L["up"] else POST_NONE, None,
                             dgamma=g[bn + "/gamma"], dbeta=g[bn + "/beta"])
            dc3 = self.split(dc, f)
            cin = self.filters[::-1][i - 2] if i > 1 else 4
            self.conv_wgrad(L["x3"], dc3, g[name + "/kernel"], cin, f)
            du = self.conv_dgrad(dc3, p[name + "/kernel"], cin)
        # ---- bottleneck (fp32 Dense kernels of the bf16 engine) ----
        z = lambda *s: torch.empty(*s, dtype=F32, device=self.dev)
        ddd = du.reshape(B, -1)
        dz, dmu, dlv, dh = z(B, self.latent), z(B, self.latent), z(B, self.latent), z(B, self.latent)
        ops.dense_bwd(S["z"], p["dec_dense/kernel"], None, ddd, g["dec_dense/kernel"], g["dec_dense/bias"], x2=S["cond"], dx1=dz)
        ops.reparam_bwd(dz, S["mu"], S["lv"], S["eps"], self.beta / B, dmu, dlv)
        ops.dense_bwd(S["h"], p["z_mean/kernel"], None, dmu, g["z_mean/kernel"], g["z_mean/bias"], dx1=dh)
        ops.dense_bwd(S["h"], p["z_log_var/kernel"], None, dlv, g["z_log_var/kernel"], g["z_log_var/bias"], dx1=dh,
                      accumulate_dx=True)
        e5 = S["enc_conv5"]["e5"]
        de5 = z(B, e5.numel() // B)
        ops.dense_bwd(e5.view(B, -1), p["enc_dense/kernel"], S["h"], dh, g["enc_dense/kernel"], g["enc_dense/bias"], act=ACT_RELU,
                      dx1=de5)
        dc_e5 = torch.empty_like(e5)
        ops.act_bwd_f32(de5, e5, ACT_LEAKY, dc_e5)
        # ---- encoder ----
        self.col_sums(dc_e5, 4, g["enc_conv5/bias"])
        dc3 = self.split(dc_e5, 4, 16)
        f4 = self.filters[-1]
        self.conv_wgrad(S["enc_conv5"]["x3"], dc3, g["enc_conv5/kernel"], f4, 4)
        dy = self.conv_dgrad(dc3, p["enc_conv5/kernel"], f4)
        for i in range(nf, 0, -1):
            name, bn, f = f"enc_conv{i}", f"enc_bn{i}", self.filters[i - 1]
            L = S[name]
            dc = self.bn_bwd(dy, L["c"], f, L["st"], ACT_LEAKY, POST_POOL2, L["idx"], dgamma=g[bn + "/gamma"],
                             dbeta=g[bn + "/beta"])
            dc3 = self.split(dc, f)
            if i > 1:
                cin = self.filters[i - 2]
                self.conv_wgrad(L["x3"], dc3, g[name + "/kernel"], cin, f)
                dy = self.conv_dgrad(dc3, p[name + "/kernel"], cin)
            else:
                self.conv_wgrad(L["x3"], dc3, g[name + "/kernel"], 4 + 4 * self.ncond, f, fold=(4, 4, self.ncond))

    def train_step(self, M, cond, eps, update=True):
        """-> [loss, pm, mse, kld] device tensor; gradients in vae_params.g; `update`: apply Keras-Adam."""
        self.vp.grad.zero_()
        m = self.forward(M, cond, eps, training=True)
        self.backward()
        if update:
            ops.adam_keras_step(self.vp.theta, self.vp.grad, self.vp.adam_m, self.vp.adam_v, self.vp.adam_state, self.lr)
        return m


class UNetForwardX3(_X3Base):
    def __init__(self, batch, d=32, channels=4, classes=95, device="cuda", params: ParamStore = None, fmt=None):
        super().__init__(device, fmt)
        assert channels == 4
        self.B, self.d, self.classes, self.pp = batch, d, classes, params
        self.taps, self.saved = {}, {}

    def forward(self, X, training=False):
        """unet_3d_multiclass (unet.py:272-355) -> fp32 logits [B,d,d,d,96] (sigmoid logit in column `classes`)."""
        p, B, d, dev = self.pp.p, self.B, self.d, self.dev
        X = X.to(dev, F32).contiguous()
        S = self.saved
        S.clear()
        x3 = torch.zeros(B, d, d, d, 48, dtype=BF16, device=dev)
        ops.pack_vae_input_split3(X, None, None, x3, fmt=self.fmt)
        cat = {k: torch.zeros(B, d >> lvl, d >> lvl, d >> lvl, 3 * c, dtype=BF16, device=dev) for k, (lvl, c) in CAT_CH.items()}
        y3, p3 = {}, {}
        for L in UNET_PLAN:
            n, cout, src = L["n"], L["cout"], L["src"]
            xin = (x3 if src == "x" else p3[src[5:]] if src.startswith("pool:") else cat[src[4:]] if src.startswith("cat:")
                   else y3[src])
            cin = 4 if n == "c1" else L["cin"]
            a = self.conv(xin, p[n + "/kernel"], p[n + "/bias"], cout, act=ACT_RELU, cin_pad=pad16(cin))
            self.taps[n] = a
            st = self.bn_coeffs(a, cout, p[f"bn_{n}/gamma"], p[f"bn_{n}/beta"], p[f"bn_{n}/moving_mean"],
                                p[f"bn_{n}/moving_variance"], training)
            rec = S[n] = dict(x3=xin, a=a, st=st, idx=None)
            if "up" in L:
                buf, off = L["up"]
                self.bn_split(a, cout, st, ACT_NONE, POST_UP2, y3=cat[buf], ctot=CAT_CH[buf][1], coff=off)
            elif "cat" in L:
                buf, off = L["cat"]
                self.bn_split(a, cout, st, ACT_NONE, POST_NONE, y3=cat[buf], ctot=CAT_CH[buf][1], coff=off)
            else:
                y3[n] = self.bn_split(a, cout, st, ACT_NONE, POST_NONE)
            if L.get("pool"):
                p3[n], rec["idx"] = self.bn_split(a, cout, st, ACT_NONE, POST_POOL2, want_idx=True)
        nh = pad16(self.classes + 1)
        wh = torch.cat([p["soft/kernel"], p["sig/kernel"]], dim=-1).contiguous()  # (1,1,1,128,classes+1)
        bh = torch.zeros(nh, dtype=F32, device=dev)
        bh[: self.classes] = p["soft/bias"]
        bh[self.classes] = p["sig/bias"][0]
        logits = self.conv(y3["c18"], wh, bh, nh)
        S["heads"] = dict(x3=y3["c18"], wh=wh)
        self.taps["soft_logits"], self.taps["sig_logit"] = logits[..., : self.classes], logits[..., self.classes:self.classes + 1]
        return logits

    def predict(self, X):
        """-> (logits fp32 [B,d,d,d,96], argmax uint8, sigmoid prob fp32) — learning phase 0 (generate.py:220-225)."""
        logits = self.forward(X, training=False)
        B, d, dev = self.B, self.d, self.dev
        argmax = torch.empty(B, d, d, d, dtype=torch.uint8, device=dev)
        sigp = torch.empty(B, d, d, d, dtype=F32, device=dev)
        ops.heads_predict(logits, self.classes, 0.8, argmax=argmax, sig_prob=sigp)
        return logits, argmax, sigp


class UNetTrainX3(UNetForwardX3):
    """One fp32-class U-Net train_on_batch (unet.py:252-259, 357-375): forward with batch statistics, weighted CCE + BCE,
    backward through skips / pools / upsamplings into params.g, Keras-Adam."""

    def __init__(self, *a, lr=1e-6, class_weight=None, **kw):
        kw["fmt"] = 0
        super().__init__(*a, **kw)
        self.lr = float(lr)
        cw = torch.full((self.classes,), float(self.classes)) if class_weight is None else torch.as_tensor(class_weight, dtype=F32)
        self.class_w = cw.to(self.dev).float().contiguous()

    def train_step(self, X, species, update=True):
        """-> metrics [loss, soft, sig, f1_m, wr_m] (device tensor); gradients in params.g."""
        p, g, B, d, dev, C1 = self.pp.p, self.pp.g, self.B, self.d, self.dev, self.classes
        self.pp.grad.zero_()
        logits = self.forward(X, training=True)
        S = self.saved
        species = species.to(dev).to(torch.uint8).reshape(B, d, d, d).contiguous()
        Mv = B * d ** 3
        nh = logits.shape[-1]
        npart = ops.heads_loss_nparts(Mv)
        partials = torch.zeros(npart, 6, dtype=F64, device=dev)
        dl = torch.zeros(B, d, d, d, nh, dtype=F32, device=dev)
        self.argmax = torch.empty(B, d, d, d, dtype=torch.uint8, device=dev)
        ops.heads_loss(logits, C1, species, self.class_w, 1.0 / Mv, partials, argmax_out=self.argmax, dlogits=dl)
        self.metrics = torch.zeros(5, dtype=F32, device=dev)
        ops.heads_loss_finalize(partials, float(Mv), self.metrics)
        # ---- heads ----
        dl3 = self.split(dl, nh, nh)
        H = S["heads"]
        dwcat = self.conv_wgrad(H["x3"], dl3, None, 128, nh, k1=True)
        colsum = self.col_sums(dl, nh, torch.empty(nh, dtype=F32, device=dev))
        ops.unpack_heads_grad(dwcat, colsum, C1, g["soft/kernel"], g["sig/kernel"], g["soft/bias"], g["sig/bias"])
        grads = {"c18": dict(dy=self.conv_dgrad(dl3, H["wh"], 128))}
        dcat = {k: torch.zeros(B, d >> lvl, d >> lvl, d >> lvl, c, dtype=F32, device=dev) for k, (lvl, c) in CAT_CH.items()}
        # ---- blocks, last to first ----
        for L in reversed(UNET_PLAN):
            n, C = L["n"], L["cout"]
            R = S[n]
            if "up" in L:
                buf, off = L["up"]
                dy, post, idx, dy2 = dcat[buf][..., off:off + C], POST_UP2, None, None
            elif "cat" in L:
                buf, off = L["cat"]
                dy, post, idx, dy2 = grads[n]["dp"], POST_POOL2, R["idx"], dcat[buf][..., off:off + C]
            else:
                dy, post, idx, dy2 = grads[n]["dy"], POST_NONE, None, None
            dc = self.bn_bwd(dy, R["a"], C, R["st"], ACT_NONE, post, idx, pre_relu=True, dy2=dy2, dgamma=g[f"bn_{n}/gamma"],
                             dbeta=g[f"bn_{n}/beta"])
            self.col_sums(dc, C, g[n + "/bias"])
            dc3 = self.split(dc, C)
            cin = 4 if n == "c1" else L["cin"]
            self.conv_wgrad(R["x3"], dc3, g[n + "/kernel"], cin, C)
            if n == "c1":
                continue
            dx = self.conv_dgrad(dc3, p[n + "/kernel"], cin)
            src = L["src"]
            if src.startswith("pool:"):
                grads.setdefault(src[5:], {})["dp"] = dx
            elif src.startswith("cat:"):
                dcat[src[4:]] = dx
            else:
                grads.setdefault(src, {})["dy"] = dx
        if update:
            ops.adam_keras_step(self.pp.theta, self.pp.grad, self.pp.adam_m, self.pp.adam_v, self.pp.adam_state, self.lr)
        return self.metrics
