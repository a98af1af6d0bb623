"""Step executors: the sequence of C-ABI kernel launches that make up one forward / backward / update
of the reference's two graphs, over pre-allocated device buffers (so a whole step can be captured in a
CUDA graph and replayed).

    VAEEngine  — conditional DFC-VAE (vae/lattice_vae.py:127-158, 160-270, 296-310) trained against the
                 frozen U-Net prefix c1..c10 used as perceptual model (lattice_vae.py:257-270)

Every FLOP runs in libicsg3d.so (icsg3d_b200/ops.py); this file only orders launches and owns buffers.
Data-parallel mode: batch sharded over ranks; BatchNorm statistic sums (forward and backward)
and the flat gradient buffer are summed INSIDE the consuming kernels over NVLink peer (symmetric) memory
(PeerBN below; csrc/bn.cu::bn_reduce_allreduce_kernel, csrc/misc.cu::adam_allreduce_kernel), so the whole
data-parallel step is one CUDA graph without a NCCL call; ICSG3D_DP_PEER=0 falls back to NCCL all-reduces between
graph segments.  Nothing else is exchanged.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, POST_NONE, POST_POOL2, POST_UP2, pad16
from .params import LATENT, VAE_FILTERS, ParamStore, unet_specs, vae_specs

BF16 = torch.bfloat16
F32 = torch.float32
F64 = torch.float64

PM_BLOCKS = [  # U-Net prefix evaluated by the perceptual loss: (name, cin, cout, level, pool_after, dfc_tap)
    ("c1", 4, 32, 0, False, False), ("c2", 32, 64, 0, True, True), ("c3", 64, 64, 1, False, False),
    ("c4", 64, 128, 1, True, True), ("c5", 128, 128, 2, False, False), ("c6", 128, 256, 2, True, True),
    ("c9", 256, 512, 3, False, False), ("c10", 512, 512, 3, False, True),
]


class Dist:
    """Thin view of torch.distributed for the two collectives the path needs."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)


class PeerBN:
    """Symmetric (NVLink peer-mapped) buffer + step epoch for the in-kernel BatchNorm-statistics all-reduce
    (csrc/bn.cu::bn_reduce_allreduce_kernel).  One per engine and process group."""

    NSLOTS, CMAX = 64, 512

    def __init__(self, dist: "Dist", dev, n_grad=0):
        import torch.distributed as td
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.world, dist.rank
        grp = dist.group if dist.group is not None else td.group.WORLD
        nbytes = ops.bn_allreduce_buffer_bytes(self.world, self.NSLOTS, self.CMAX)
        self.buf = symm.empty(nbytes // 8, dtype=torch.int64, device=dev)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, grp)
        self.peers = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        # second symmetric buffer: the flat gradient of every rank (all-reduce fused into the Adam kernel)
        self.grad_peers = None
        if n_grad > 0:
            gbytes = ops.adam_allreduce_buffer_bytes(self.world, n_grad)
            self.gbuf = symm.empty((gbytes + 7) // 8, dtype=torch.int64, device=dev)
            self.gbuf.zero_()
            self.ghdl = symm.rendezvous(self.gbuf, grp)
            self.grad_peers = torch.tensor([int(p) for p in self.ghdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self._slots = {}
        torch.cuda.synchronize()
        td.barrier(group=dist.group)  # every rank's buffer is zeroed before anybody pushes into it

    def slot(self, key):
        s = self._slots.setdefault(key, len(self._slots))
        assert s < self.NSLOTS, "PeerBN: out of all-reduce slots"
        return s

    def tick(self):
        self.epoch.add_(1)

    def args(self, key):
        return dict(peers=self.peers, world=self.world, rank=self.rank, slot=self.slot(key), nslots=self.NSLOTS,
                    cmax=self.CMAX, epoch=self.epoch)


class _BN:
    """Per-layer BatchNorm scratch: batch mean/rstd, folded scale/shift, fp64 statistic sums."""

    def __init__(self, C, dev):
        self.C = C
        self.mean = torch.zeros(C, dtype=F32, device=dev)
        self.rstd = torch.zeros(C, dtype=F32, device=dev)
        self.scale = torch.zeros(C, dtype=F32, device=dev)
        self.shift = torch.zeros(C, dtype=F32, device=dev)
        self.sums = torch.zeros(2 * C, dtype=F64, device=dev)
        self.bsums = torch.zeros(2 * C, dtype=F64, device=dev)
        self.bsums_g = torch.zeros(2 * C, dtype=F64, device=dev)


class _Ctx:
    """Scratch shared by the launches of one stream."""

    def __init__(self, dev, max_partials=148 * 4 * 2 * 1024, max_dw=27 * 512 * 512, conv_ws_bytes=0):
        self.partials = torch.zeros(max_partials, dtype=F64, device=dev)
        self.dw_pad = torch.zeros(max_dw, dtype=F32, device=dev)
        # split-K partials of the per-tap conv kernel (4^3 / 2^3 layers); one buffer per stream
        self.conv_ws = torch.empty(max(conv_ws_bytes, 16), dtype=torch.uint8, device=dev)


class VAEEngine:
    def __init__(self, batch, d=32, channels=4, ncond=10, latent=LATENT, filters=VAE_FILTERS, device="cuda",
                 vae_params: ParamStore | None = None, pm_params: ParamStore | None = None, alpha=0.5, beta=3e-4,
                 pm_layer_weights=(1.0, 1.0, 1.0, 1.0), lr=5e-4, dist: Dist | None = None, seed=1):
        assert channels == 4, "the reference hard-wires 4 input channels (lattice_vae.py:90)"
        assert d % 16 == 0 and d & (d - 1) == 0, "grid edge must be a power of two >= 16"
        self.B, self.d, self.ncond, self.latent, self.filters = batch, d, ncond, latent, list(filters)
        self.dev = torch.device(device)
        self.alpha, self.beta, self.lr = float(alpha), float(beta), float(lr)
        self.pm_w = [float(w) for w in pm_layer_weights]
        self.dist = dist
        self.world = dist.world if dist else 1
        dev = self.dev
        self.vp = vae_params or ParamStore(vae_specs(channels, ncond, d, latent, filters), dev).init(seed)
        self.pp = pm_params or ParamStore(unet_specs(channels), dev, with_grads=False, with_adam=False).init(seed + 1)
        ws_bytes = self._conv_ws_bytes(batch, d, filters)
        self.ctx = _Ctx(dev, conv_ws_bytes=ws_bytes)
        # pm(x) does not depend on the encoder/decoder: it runs on a side stream (forked/joined inside the captured
        # graph) so that its large convs overlap the small, latency-bound VAE kernels.  Own scratch per stream.
        self.ctx2 = _Ctx(dev, max_dw=16, conv_ws_bytes=ws_bytes)
        # filter-gradient scratch: owned by this engine (all its wgrad launches are ordered on one stream), sized once for
        # the largest layer and never reallocated — its address is baked into the captured step graph
        self.wg_ws = torch.empty(max(self._wgrad_ws_bytes(batch, d, filters), 16), dtype=torch.uint8, device=dev)
        # Data parallel: BatchNorm statistic sums are exchanged INSIDE the finalize kernels over NVLink peer memory
        # (PeerBN); only the flat gradient goes through NCCL.  ICSG3D_DP_PEER=0 (or no symmetric memory) falls back to
        # NCCL all-reduces of the sums, which also forces a single stream (collectives split the captured graph).
        self.peer = None
        if self.world > 1 and os.environ.get("ICSG3D_DP_PEER", "1") != "0":
            try:
                self.peer = PeerBN(dist, self.dev, n_grad=self._vp_numel())
            except Exception as e:  # noqa: BLE001
                import warnings
                warnings.warn(f"icsg3d: symmetric-memory BatchNorm all-reduce unavailable ({e!r}); using NCCL")
        self.overlap_pm = self.world == 1 or self.peer is not None
        self._side = None
        B = batch
        z = lambda *s, dt=BF16: torch.zeros(*s, dtype=dt, device=dev)
        # Encoder FORWARD on fp32-class split operands (bf16 pairs hi + lo on the ordinary tcgen05 kernels, csrc/split3.cu):
        # the KLD metric is a cancelling sum over mu / log-var whose bf16 error is systematic across the batch (every
        # rounding point of the encoder — input, weights, conv outputs, pooled outputs — moves it by ~4e-4 relative with a
        # random sign; measured 1.6e-3 at batch 32), so it is the one part of the step that needs more than bf16 to meet
        # the 1e-3 loss bar.  The encoder is 3 % of the step's executed FLOPs; its backward stays bf16 on the hi parts.
        # ICSG3D_ENC_FP32=0 restores the plain bf16 encoder (A/B timing).
        self.enc_x3 = os.environ.get("ICSG3D_ENC_FP32", "1") != "0"
        kx = 3 if self.enc_x3 else 1

        # ---- static inputs ----
        self.M = z(B, d, d, d, 4, dt=F32)
        self.cond = z(B, ncond, dt=F32)
        self.eps = z(B, latent, dt=F32)
        # split encoder input: lean 32-channel layout (the one-hot condition needs no lo part; csrc/split3.cu), whose first
        # 16 channels start with the plain bf16 operand (M_hi, cond) that the bf16 filter gradient of enc_conv1 reads in place
        self.enc_lean = self.enc_x3 and ncond <= 10 and os.environ.get("ICSG3D_ENC_LEAN", "1") != "0"
        self.xe3 = z(B, d, d, d, 32 if self.enc_lean else 16 * kx)
        self.xe = self.xe3[..., :16]
        self.xp = z(B, d, d, d, 16)

        # ---- encoder ----
        self.enc = []
        cin_pad, D = 16, d
        cdt = F32 if self.enc_x3 else BF16   # conv outputs / incoming gradients of the split layers stay fp32
        for i, f in enumerate(self.filters, 1):
            y3 = z(B, D // 2, D // 2, D // 2, f * kx)
            L = dict(name=f"enc_conv{i}", bn=f"enc_bn{i}", cin_pad=cin_pad, cout=f, D=D,
                     c=z(B, D, D, D, f, dt=cdt), y3=y3, y=y3[..., :f], idx=z(B, D // 2, D // 2, D // 2, f, dt=torch.uint8),
                     dc=z(B, D, D, D, f), dy=z(B, D // 2, D // 2, D // 2, f, dt=cdt), bns=_BN(f, dev),
                     wf=z(27, f, 32 if (i == 1 and self.enc_lean) else cin_pad * kx), wd=z(27, cin_pad, f) if i > 1 else None)
            self.enc.append(L)
            cin_pad, D = f, D // 2
        self.e_s = D  # spatial edge at enc_conv5 (d/16)
        c4 = self.filters[-1]
        self.e5 = z(B, D, D, D, 4, dt=F32)
        self.e5_wf, self.e5_wd = z(27, 16, c4 * kx), z(27, c4, 16)
        self.de5 = z(B, D * D * D * 4, dt=F32)
        self.dc_e5 = z(B, D, D, D, 16)
        # ---- bottleneck ----
        self.h, self.mu, self.lv, self.z, self.kl = (z(B, latent, dt=F32), z(B, latent, dt=F32), z(B, latent, dt=F32),
                                                      z(B, latent, dt=F32), z(B, dt=F32))
        self.dh, self.dmu, self.dlv, self.dz = (z(B, latent, dt=F32), z(B, latent, dt=F32), z(B, latent, dt=F32),
                                                  z(B, latent, dt=F32))
        s0 = d // 8
        self.s0 = s0
        self.dd = z(B, s0 * s0 * s0 * 4, dt=F32)
        self.ddd = z(B, s0 * s0 * s0 * 4, dt=F32)
        self.d0 = z(B, s0, s0, s0, 16)
        self.dy_d0 = z(B, s0, s0, s0, 16)
        # ---- decoder ----
        self.dec = []
        cin_pad, S = 16, s0
        for i, f in enumerate(self.filters[::-1], 1):
            up = i < len(self.filters)
            So = S * 2 if up else S
            L = dict(name=f"dec_conv{i}", bn=f"dec_bn{i}", cin_pad=cin_pad, cout=f, D=S, up=up,
                     c=z(B, S, S, S, f), u=z(B, So, So, So, f), dc=z(B, S, S, S, f), du=z(B, So, So, So, f),
                     bns=_BN(f, dev), wf=z(27, f, cin_pad), wd=z(27, cin_pad, f))
            self.dec.append(L)
            cin_pad, S = f, So
        cl = self.filters[0]
        self.c5 = z(B, d, d, d, 4, dt=F32)
        self.xhat = z(B, d, d, d, 4, dt=F32)
        self.xhat16 = z(B, d, d, d, 16)
        self.dxhat = z(B, d, d, d, 4, dt=F32)
        self.dc5 = z(B, d, d, d, 16)
        self.out_wf, self.out_wd = z(27, 16, cl), z(27, cl, 16)
        self.bn5 = _BN(4, dev)
        # ---- perceptual model (two branches: 0 = x, 1 = x_hat) ----
        self.pm = []
        for name, cin, cout, lvl, pool, tap in PM_BLOCKS:
            D = d >> lvl
            cp = pad16(cin)
            L = dict(name=name, cin_pad=cp, cout=cout, D=D, pool=pool, tap=tap, has_bn=name != "c10",
                     wf=ops.pack_conv_w_fprop(self.pp.p[name + "/kernel"]),
                     wd=ops.pack_conv_w_dgrad(self.pp.p[name + "/kernel"]),
                     a=[z(B, D, D, D, cout), z(B, D, D, D, cout)], bnst=[_BN(cout, dev), _BN(cout, dev)])
            Do = D // 2 if pool else D
            if L["has_bn"]:
                L["y"] = [z(B, Do, Do, Do, cout), z(B, Do, Do, Do, cout)]
                L["idx"] = [z(B, Do, Do, Do, cout, dt=torch.uint8) if pool else None for _ in range(2)]
                L["dy"] = z(B, Do, Do, Do, cout)   # gradient w.r.t. y (x_hat branch only)
            L["dc"] = z(B, D, D, D, cout)
            self.pm.append(L)
        self.dxh16 = z(B, d, d, d, 16)
        # ---- losses ----
        self.n_terms = 1 + sum(1 for b in PM_BLOCKS if b[5])
        self.loss_stride = 148 * 4
        self.loss_partials = torch.zeros(self.n_terms, self.loss_stride, dtype=F64, device=dev)
        self.loss_nparts = torch.zeros(self.n_terms, dtype=torch.int32, device=dev)
        self.loss_scales = torch.zeros(self.n_terms, dtype=F64, device=dev)
        self.metrics = torch.zeros(4, dtype=F32, device=dev)
        self._loss_meta_ready = False
        self._graph = None
        self._segments = None
        self.use_graph = False
        self._pack_table = None
        self._wg_side = None
        self.overlap_wgrad = self.world == 1 or self.peer is not None  # filter gradients on a side stream, under the BN-backward / dgrad chain
        # whole BatchNorm backward per layer in one cooperative launch (needs the peer-memory path when data parallel).
        # Measured slower inside the step (3.59 vs 3.35 ms: the co-resident grid is half as wide and the cooperative launch
        # does not overlap with the filter-gradient side stream), so it is opt-in: ICSG3D_FUSE_BN_BWD=1.
        self.fuse_bn_bwd = os.environ.get("ICSG3D_FUSE_BN_BWD", "0") == "1" and (self.world == 1 or self.peer is not None)
        # DFC feature-loss sums out of the BatchNorm-backward apply pass of the tapped layers (ICSG3D_FUSE_TAP_LOSS=0: separate pass)
        self._fuse_small_ok = self.world == 1 or self.peer is not None
        self.fuse_bn_bwd_small_bytes = int(float(os.environ.get("ICSG3D_FUSE_BN_BWD_SMALL_MB", "0")) * (1 << 20))
        self.fuse_tap_loss = os.environ.get("ICSG3D_FUSE_TAP_LOSS", "1") != "0"
        self._defer_taps = False
        self.fuse_stats = True  # BatchNorm statistics from the conv epilogue where the streaming kernel serves the layer
        self._exp_skip_wgrad = os.environ.get("ICSG3D_EXP_SKIP_WGRAD", "0") == "1"
        self._exp_skip_pm0 = os.environ.get("ICSG3D_EXP_SKIP_PM0", "0") == "1"

    # ------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _conv_ws_bytes(B, d, filters):
        """Largest split-K workspace any conv of the step asks for (fprop and dgrad operand shapes)."""
        shapes = [(d >> lvl, pad16(cin), cout) for _, cin, cout, lvl, _, _ in PM_BLOCKS]
        f = list(filters)
        D = d
        cin = 16
        for c in f:
            shapes.append((D, cin, c))
            shapes.append((D, 3 * cin, c))  # split-operand encoder forward (3x the K extent)
            shapes.append((D, 2 * cin, c))  # lean split layout of the first layer
            cin, D = c, D // 2
        shapes.append((D, f[-1], 16))
        shapes.append((D, 3 * f[-1], 16))
        S = d // 8
        cin = 16
        for i, c in enumerate(f[::-1]):
            shapes.append((S, cin, c))
            cin, S = c, (S * 2 if i < len(f) - 1 else S)
        need = 0
        for D_, ci, co in shapes:
            need = max(need, ops.conv3d_k3_workspace_bytes(B, D_, ci, co), ops.conv3d_k3_workspace_bytes(B, D_, co, ci))
        return need

    @staticmethod
    def _wgrad_ws_bytes(B, d, filters):
        """Largest filter-gradient workspace any layer of the step asks for."""
        f = list(filters)
        shapes, D, cin = [], d, 16
        for c in f:  # encoder: (D, cin_pad, cout)
            shapes.append((D, cin, c))
            cin, D = c, D // 2
        shapes.append((D, f[-1], 16))  # enc_conv5
        S, cin = d // 8, 16
        for i, c in enumerate(f[::-1]):
            shapes.append((S, cin, c))
            cin, S = c, (S * 2 if i < len(f) - 1 else S)
        shapes.append((S, f[0], 16))  # decoder_output
        return max(ops.conv3d_k3_wgrad_workspace_bytes(B, D_, ci, co) for D_, ci, co in shapes)

    def _conv(self, x, w, bias, ctx=None, **kw):
        """Conv3D through the dispatcher with this stream's split-K workspace (used by the 4^3 / 2^3 layers)."""
        return ops.conv3d_k3(x, w, bias, ws=(ctx or self.ctx).conv_ws, **kw)

    def _conv_bn(self, x, wf, bias, out, C, training, ctx=None, **kw):
        """Conv3D whose output feeds a BatchNorm: when the layer is served by the plane-streaming kernel the batch
        statistics come out of the conv epilogue (no separate read pass).  Returns the partials view or None."""
        part = None
        if training and self.fuse_stats and out.shape[-1] == C:  # bf16 or fp32 output (statistics of the stored values)
            n = ops.conv3d_k3_stats_parts(x, wf)
            if n > 0:
                part = (ctx or self.ctx).partials[: n * 2 * C].view(n, 2, C)
        self._conv(x, wf, bias, ctx=ctx, out=out, stats=part, **kw)
        return part

    def _bn_fwd(self, x, C, st: _BN, gamma, beta, mm, mv, training, act, post, y=None, y32=None, idx=None, ctx=None,
                part=None, y_split=None):
        if training:
            rows = x.numel() // x.shape[-1]
            if part is None:
                n = ops.bn_nparts(rows, C, x.dtype)
                part = (ctx or self.ctx).partials[: n * 2 * C].view(n, 2, C)
                ops.bn_stats(x, C, part)
            if self.peer is not None:
                ops.bn_reduce_allreduce_finalize(part, float(rows * self.world), gamma, beta, st.sums, st.mean, st.rstd,
                                                 st.scale, st.shift, moving_mean=mm, moving_var=mv,
                                                 **self.peer.args((id(st), "fwd")))
            elif self.world > 1:
                ops.bn_reduce_partials(part, st.sums)
                self.dist.all_reduce_sum(st.sums)
                ops.bn_finalize(st.sums, float(rows * self.world), gamma, beta, st.mean, st.rstd, st.scale, st.shift, mm, mv)
            else:
                ops.bn_reduce_finalize(part, float(rows), gamma, beta, st.sums, st.mean, st.rstd, st.scale, st.shift, mm, mv)
        else:
            ops.bn_inference_coeffs(gamma, beta, mm, mv, st.scale, st.shift)
        if y_split is not None:  # fp32 conv output -> [hi | lo | hi] bf16 pairs for the next split conv
            ops.bn_apply_fwd_split3(x, C, st.scale, st.shift, act, post, y_split, C, 0, pool_idx=idx, fmt=0)
        else:
            ops.bn_apply_fwd(x, C, st.scale, st.shift, act, post, y=y, y32=y32, pool_idx=idx)

    def _bn_bwd(self, dy, x, C, st: _BN, act, post, idx, dx, pre_relu=False, tap_other=None, tap_coef=0.0, dgamma=None,
                dbeta=None, tap_sq=None):
        rows = x.numel() // x.shape[-1]
        assert tap_sq is None or not self.fuse_bn_bwd
        # optional (ICSG3D_FUSE_BN_BWD_SMALL_MB > 0; measured 3.04 vs 2.99 ms/step at 6 MB, so off by default): layers of a
        # few MB take the one-launch cooperative kernel
        small = tap_sq is None and self._fuse_small_ok and x.numel() * x.element_size() <= self.fuse_bn_bwd_small_bytes
        if self.fuse_bn_bwd or small:
            # one cooperative launch: partial sums -> grid barrier -> fixed-order reduction (+ peer-memory all-reduce in
            # data-parallel mode) -> grid barrier -> dx; the second read of dy / x comes from L2 where the layer fits
            n = ops.bn_bwd_fused_nparts(C, x.dtype)
            part = self.ctx.partials[: n * 2 * C].view(n, 2, C)
            kw = self.peer.args((id(st), "bwd")) if self.peer is not None else {}
            ops.bn_bwd_fused(dy, x, C, st.mean, st.rstd, st.scale, st.shift, act, post, idx, part, st.bsums_g,
                             float(rows * self.world), dx, dgamma=dgamma, dbeta=dbeta, pre_relu=pre_relu, tap_other=tap_other,
                             tap_coef=tap_coef, **kw)
            return
        n = ops.bn_bwd_nparts(x, C, post)
        part = self.ctx.partials[: n * 2 * C].view(n, 2, C)
        ops.bn_bwd_reduce(dy, x, C, st.mean, st.rstd, st.scale, st.shift, act, post, idx, part)
        if self.peer is not None:
            ops.bn_reduce_allreduce_grads(part, st.bsums_g, dgamma=dgamma, dbeta=dbeta, **self.peer.args((id(st), "bwd")))
            ops.bn_bwd_apply(dy, x, C, st.mean, st.rstd, st.scale, st.shift, act, post, idx, st.bsums_g,
                             float(rows * self.world), dx, pre_relu=pre_relu, tap_other=tap_other, tap_coef=tap_coef,
                             tap_sq=tap_sq)
            return
        ops.bn_reduce_grads(part, st.bsums, dgamma, dbeta)  # local sums: the gradient all-reduce adds the ranks
        sums = st.bsums
        if self.world > 1:
            st.bsums_g.copy_(st.bsums)
            self.dist.all_reduce_sum(st.bsums_g)
            sums = st.bsums_g
        ops.bn_bwd_apply(dy, x, C, st.mean, st.rstd, st.scale, st.shift, act, post, idx, sums, float(rows * self.world), dx,
                         pre_relu=pre_relu, tap_other=tap_other, tap_coef=tap_coef, tap_sq=tap_sq)

    def _wgrad(self, x, dy, name, cin, cout, cin_pad, cout_pad, fold=None):
        """dW of conv `name` into the flat gradient buffer (Keras layout).  Off the critical path (only Adam needs it):
        launched on a side stream that forks here and joins before the optimiser step."""
        if self._exp_skip_wgrad:  # TIMING EXPERIMENT ONLY (ICSG3D_EXP_SKIP_WGRAD=1): how much of the side-stream work is hidden
            return
        if not self.overlap_wgrad:
            return self._wgrad_now(x, dy, name, cin, cout, cin_pad, cout_pad, fold)
        main = torch.cuda.current_stream()
        if self._wg_side is None:
            self._wg_side = torch.cuda.Stream()
        self._wg_side.wait_stream(main)
        with torch.cuda.stream(self._wg_side):
            self._wgrad_now(x, dy, name, cin, cout, cin_pad, cout_pad, fold)
        self._wg_pending = True

    def _wgrad_join(self):
        if self.overlap_wgrad and getattr(self, "_wg_pending", False):
            torch.cuda.current_stream().wait_stream(self._wg_side)
            self._wg_pending = False

    def _wgrad_now(self, x, dy, name, cin, cout, cin_pad, cout_pad, fold=None):
        g = self.vp.g[name + "/kernel"]
        if cin == cin_pad and cout == cout_pad and fold is None:
            ops.conv3d_k3_wgrad(x, dy, cin=cin_pad, cout=cout_pad, out=g.view(27, cin, cout), tag=name + ".wgrad",
                                ws=self.wg_ws)
            return
        scratch = self.ctx.dw_pad[: 27 * cin_pad * cout_pad].view(27, cin_pad, cout_pad)
        ops.conv3d_k3_wgrad(x, dy, cin=cin_pad, cout=cout_pad, out=scratch, tag=name + ".wgrad", nominal=(cin, cout),
                            ws=self.wg_ws)
        if fold is None:
            ops.unpack_conv_dw(scratch, cin, cout, out=g)
        else:
            ops.unpack_conv_dw(scratch, cin, cout, cin_lead=fold[0], fold=fold[1], fold_c=fold[2], out=g)

    def pack_weights(self):
        """fp32 master weights -> bf16 GEMM operand layouts (after every optimiser step): one launch for all 19 packs."""
        if self._pack_table is None:
            p = self.vp.p
            jobs = []
            fm = 2 if self.enc_x3 else 0  # pack mode of the encoder's fprop operands (2 = bf16-pair split)
            for i, L in enumerate(self.enc):
                if i == 0:
                    jobs.append((p[L["name"] + "/kernel"], L["wf"], 4 if self.enc_lean else fm, 4, 4, self.ncond))
                else:
                    jobs.append((p[L["name"] + "/kernel"], L["wf"], fm, 0, 1, 0))
                    jobs.append((p[L["name"] + "/kernel"], L["wd"], 1, 0, 1, 0))
            jobs.append((p["enc_conv5/kernel"], self.e5_wf, fm, 0, 1, 0))
            jobs.append((p["enc_conv5/kernel"], self.e5_wd, 1, 0, 1, 0))
            for L in self.dec:
                jobs.append((p[L["name"] + "/kernel"], L["wf"], 0, 0, 1, 0))
                jobs.append((p[L["name"] + "/kernel"], L["wd"], 1, 0, 1, 0))
            jobs.append((p["decoder_output/kernel"], self.out_wf, 0, 0, 1, 0))
            jobs.append((p["decoder_output/kernel"], self.out_wd, 1, 0, 1, 0))
            self._pack_table = ops.pack_jobs_table(jobs, self.dev)
        ops.pack_conv_w_batch(self._pack_table)

    def pack_inputs(self):
        """fp32 batch + one-hot condition -> encoder operand (bf16, or [hi | lo | hi] bf16 pairs for the split encoder) and
        the perceptual U-Net's bf16 operand."""
        if self.enc_lean:
            ops.pack_vae_input_lean(self.M, self.cond, self.xe3, self.xp)
        elif self.enc_x3:
            ops.pack_vae_input_mixed(self.M, self.cond, self.xe3, self.xp, fmt=0)
        else:
            ops.pack_vae_input(self.M, self.cond, self.xe3, self.xp)

    def repack_pm(self):
        for L in self.pm:
            ops.pack_conv_w_fprop(self.pp.p[L["name"] + "/kernel"], out=L["wf"])
            ops.pack_conv_w_dgrad(self.pp.p[L["name"] + "/kernel"], out=L["wd"])

    # ------------------------------------------------------------------------------------------
    # forward pieces
    # ------------------------------------------------------------------------------------------
    def encode(self, training):
        """build_encoder (lattice_vae.py:160-195) on self.xe / self.eps -> self.mu, self.lv, self.z."""
        p = self.vp.p
        x = self.xe3
        for L in self.enc:
            nominal = (4 + 4 * self.ncond, L["cout"]) if L is self.enc[0] else (L["cin_pad"], L["cout"])
            part = self._conv_bn(x, L["wf"], p[L["name"] + "/bias"], L["c"], L["cout"], training, tag=L["name"] + ".fprop",
                                 nominal=nominal)
            bn = L["bn"]
            self._bn_fwd(L["c"], L["cout"], L["bns"], p[bn + "/gamma"], p[bn + "/beta"],
                         p[bn + "/moving_mean"], p[bn + "/moving_variance"], training, ACT_LEAKY, POST_POOL2, y=L["y"],
                         idx=L["idx"], part=part, y_split=L["y3"] if self.enc_x3 else None)
            x = L["y3"]
        self._conv(x, self.e5_wf, p["enc_conv5/bias"], out=self.e5, n_store=4, act=ACT_LEAKY, tag="enc_conv5.fprop",
                      nominal=(self.filters[-1], 4))
        ops.dense_fwd(self.e5.view(self.B, -1), p["enc_dense/kernel"], p["enc_dense/bias"], self.h, act=ACT_RELU)
        ops.dense_fwd(self.h, p["z_mean/kernel"], p["z_mean/bias"], self.mu)
        ops.dense_fwd(self.h, p["z_log_var/kernel"], p["z_log_var/bias"], self.lv)
        ops.reparam_fwd(self.mu, self.lv, self.eps, self.z, self.kl)

    def decode(self, training):
        """build_decoder (lattice_vae.py:197-230) on self.z / self.cond -> self.xhat (fp32) and self.xhat16 (bf16)."""
        p = self.vp.p
        ops.dense_fwd(self.z, p["dec_dense/kernel"], p["dec_dense/bias"], self.dd, x2=self.cond)
        ops.f32_to_bf16_rows(self.dd, 4, self.d0)
        x = self.d0
        for L in self.dec:
            part = self._conv_bn(x, L["wf"], p[L["name"] + "/bias"], L["c"], L["cout"], training, tag=L["name"] + ".fprop",
                                 nominal=(4, L["cout"]) if L is self.dec[0] else None)
            bn = L["bn"]
            self._bn_fwd(L["c"], L["cout"], L["bns"], p[bn + "/gamma"], p[bn + "/beta"], p[bn + "/moving_mean"],
                         p[bn + "/moving_variance"], training, ACT_LEAKY, POST_UP2 if L["up"] else POST_NONE, y=L["u"],
                         part=part)
            x = L["u"]
        self._conv(x, self.out_wf, p["decoder_output/bias"], out=self.c5, n_store=4, tag="decoder_output.fprop",
                      nominal=(self.filters[0], 4))
        self._bn_fwd(self.c5, 4, self.bn5, p["dec_bn5/gamma"], p["dec_bn5/beta"], p["dec_bn5/moving_mean"],
                     p["dec_bn5/moving_variance"], training, ACT_RELU, POST_NONE, y=self.xhat16, y32=self.xhat)

    def pm_forward(self, branch, training, ctx=None):
        """U-Net prefix c1..c10 (unet.py:276-306) on x (branch 0) or x_hat (branch 1); BN follows the learning
        phase and its moving averages are never updated from here."""
        p = self.pp.p
        x = self.xp if branch == 0 else self.xhat16
        for L in self.pm:
            n = L["name"]
            kw = dict(act=ACT_RELU, tag=f"pm{branch}.{n}.fprop", nominal=(4, L["cout"]) if n == "c1" else None)
            if not L["has_bn"]:
                self._conv(x, L["wf"], p[n + "/bias"], ctx=ctx, out=L["a"][branch], **kw)
                break
            part = self._conv_bn(x, L["wf"], p[n + "/bias"], L["a"][branch], L["cout"], training, ctx=ctx, **kw)
            bn = "bn_" + n
            self._bn_fwd(L["a"][branch], L["cout"], L["bnst"][branch], p[bn + "/gamma"], p[bn + "/beta"],
                         p[bn + "/moving_mean"] if not training else None,
                         p[bn + "/moving_variance"] if not training else None, training, ACT_NONE,
                         POST_POOL2 if L["pool"] else POST_NONE, y=L["y"][branch], idx=L["idx"][branch], ctx=ctx, part=part)
            x = L["y"][branch]

    def _loss_meta(self):
        if self._loss_meta_ready:
            return
        nparts, scales = [], []
        n_mse = self.M.numel()
        nparts.append(ops.sqdiff_nparts(n_mse))
        scales.append(1.0 / n_mse)  # local batch means; metrics_host() averages the ranks
        k = 0
        for L in self.pm:
            if L["tap"]:
                n = L["a"][0].numel()
                nparts.append(ops.sqdiff_nparts(n))
                scales.append(self.pm_w[k] / n)
                k += 1
        assert max(nparts) <= self.loss_stride
        self.loss_nparts.copy_(torch.tensor(nparts, dtype=torch.int32))
        self.loss_scales.copy_(torch.tensor(scales, dtype=F64))
        self._nparts_host = nparts
        # train step: the feature-loss sums of the tapped layers that have a BatchNorm backward (all but the last) come
        # out of bn_bwd_apply (which reads both feature maps anyway) instead of a separate pass over them
        fused = list(nparts)
        k = 1
        for li, L in enumerate(self.pm):
            if L["tap"]:
                if li < len(self.pm) - 1:
                    fused[k] = ops.bn_bwd_apply_nblocks(L["a"][1], L["cout"], POST_POOL2 if L["pool"] else POST_NONE)
                k += 1
        assert max(fused) <= self.loss_stride
        self._nparts_fused_host = fused
        self.loss_nparts_fused = torch.tensor(fused, dtype=torch.int32, device=self.loss_nparts.device)
        self._loss_meta_ready = True

    def losses(self, defer_taps=False):
        """[loss, pm, mse, kld] (lattice_vae.py:241-255) into self.metrics (local batch means).  defer_taps (train step):
        the tapped layers' sums are produced by backward(), which then assembles the metrics (assemble_losses)."""
        self._loss_meta()
        ops.sqdiff_partials(self.M, self.xhat, self.loss_partials[0], self._nparts_host[0])
        k = 1
        for li, L in enumerate(self.pm):
            if L["tap"]:
                if not (defer_taps and li < len(self.pm) - 1):
                    ops.sqdiff_partials(L["a"][0], L["a"][1], self.loss_partials[k], self._nparts_host[k])
                k += 1
        if not defer_taps:
            self.assemble_losses(False)

    def assemble_losses(self, fused):
        ops.vae_loss_assemble(self.loss_partials, self.loss_nparts_fused if fused else self.loss_nparts, self.loss_scales,
                              self.kl, 1.0 / self.B, self.alpha, self.beta, self.metrics)

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def backward(self):
        p, g = self.vp.p, self.vp.g
        Bg = self.B * self.world
        # ---- perceptual branch on x_hat: dgrad only (pm is frozen) ----
        taps = [L for L in self.pm if L["tap"]]
        coef = {}
        for k, L in enumerate(taps):
            feat = L["a"][0].numel() // self.B
            coef[L["name"]] = self.alpha * self.pm_w[k] * 2.0 / (feat * Bg)
        last = self.pm[-1]
        ops.tap_grad_relu(last["a"][1], last["a"][0], coef[last["name"]], last["dc"])
        dy = self.pm[-2]["dy"]
        self._conv(last["dc"], last["wd"], None, out=dy, tag="pm1.c10.dgrad")
        tap_slot = {}
        k = 1
        for L in self.pm:
            if L["tap"]:
                tap_slot[L["name"]] = k
                k += 1
        for li in range(len(self.pm) - 2, -1, -1):
            L = self.pm[li]
            tap_sq = None
            if self._defer_taps and L["tap"]:
                ks = tap_slot[L["name"]]
                tap_sq = self.loss_partials[ks][: self._nparts_fused_host[ks]]
            self._bn_bwd(dy, L["a"][1], L["cout"], L["bnst"][1], ACT_NONE, POST_POOL2 if L["pool"] else POST_NONE,
                         L["idx"][1], L["dc"], pre_relu=True, tap_other=L["a"][0] if L["tap"] else None,
                         tap_coef=coef.get(L["name"], 0.0), tap_sq=tap_sq)
            if li > 0:
                dy = self.pm[li - 1]["dy"]
                self._conv(L["dc"], L["wd"], None, out=dy, tag=f"pm1.{L['name']}.dgrad")
            else:
                self._conv(L["dc"], L["wd"], None, out=self.dxh16, tag="pm1.c1.dgrad", nominal=(L["cout"], 4))
        if self._defer_taps:
            self.assemble_losses(True)
        # ---- decoder ----
        ops.xhat_grad(self.M, self.xhat, 2.0 / (self.M.numel() // self.B * Bg), self.dxh16, self.dxhat)
        self._bn_bwd(self.dxhat, self.c5, 4, self.bn5, ACT_RELU, POST_NONE, None, self.dc5, dgamma=g["dec_bn5/gamma"],
                     dbeta=g["dec_bn5/beta"])
        lastd = self.dec[-1]
        self._wgrad(lastd["u"], self.dc5, "decoder_output", lastd["cout"], 4, lastd["cout"], 16)
        self._conv(self.dc5, self.out_wd, None, out=lastd["du"], tag="decoder_output.dgrad", nominal=(4, lastd["cout"]))
        for li in range(len(self.dec) - 1, -1, -1):
            L = self.dec[li]
            bn = L["bn"]
            self._bn_bwd(L["du"], L["c"], L["cout"], L["bns"], ACT_LEAKY, POST_UP2 if L["up"] else POST_NONE, None, L["dc"],
                         dgamma=g[bn + "/gamma"], dbeta=g[bn + "/beta"])
            xin = self.dec[li - 1]["u"] if li > 0 else self.d0
            cin = self.dec[li - 1]["cout"] if li > 0 else 4
            self._wgrad(xin, L["dc"], L["name"], cin, L["cout"], L["cin_pad"], L["cout"])
            self._conv(L["dc"], L["wd"], None, out=self.dec[li - 1]["du"] if li > 0 else self.dy_d0,
                          tag=L["name"] + ".dgrad", nominal=(L["cout"], 4) if li == 0 else None)
        # ---- bottleneck ----
        ops.bf16_rows_to_f32(self.dy_d0, 4, self.ddd)
        ops.dense_bwd(self.z, p["dec_dense/kernel"], None, self.ddd, g["dec_dense/kernel"], g["dec_dense/bias"], x2=self.cond,
                      dx1=self.dz)
        ops.reparam_bwd(self.dz, self.mu, self.lv, self.eps, self.beta / Bg, self.dmu, self.dlv)
        ops.dense_bwd(self.h, p["z_mean/kernel"], None, self.dmu, g["z_mean/kernel"], g["z_mean/bias"], dx1=self.dh)
        ops.dense_bwd(self.h, p["z_log_var/kernel"], None, self.dlv, g["z_log_var/kernel"], g["z_log_var/bias"], dx1=self.dh,
                      accumulate_dx=True)
        ops.dense_bwd(self.e5.view(self.B, -1), p["enc_dense/kernel"], self.h, self.dh, g["enc_dense/kernel"],
                      g["enc_dense/bias"], act=ACT_RELU, dx1=self.de5)
        ops.leaky_bwd_rows(self.de5, self.e5, 4, self.dc_e5)
        # ---- encoder ----
        e4 = self.enc[-1]
        ops.bias_grad(self.dc_e5, 4, g["enc_conv5/bias"])
        self._wgrad(e4["y"], self.dc_e5, "enc_conv5", e4["cout"], 4, e4["cout"], 16)
        self._conv(self.dc_e5, self.e5_wd, None, out=e4["dy"], tag="enc_conv5.dgrad", nominal=(4, e4["cout"]))
        for li in range(len(self.enc) - 1, -1, -1):
            L = self.enc[li]
            bn = L["bn"]
            self._bn_bwd(L["dy"], L["c"], L["cout"], L["bns"], ACT_LEAKY, POST_POOL2, L["idx"], L["dc"],
                         dgamma=g[bn + "/gamma"], dbeta=g[bn + "/beta"])
            if li > 0:
                self._wgrad(self.enc[li - 1]["y"], L["dc"], L["name"], L["cin_pad"], L["cout"], L["cin_pad"], L["cout"])
                self._conv(L["dc"], L["wd"], None, out=self.enc[li - 1]["dy"], tag=L["name"] + ".dgrad")
            else:
                self._wgrad(self.xe, L["dc"], L["name"], 4 + 4 * self.ncond, L["cout"], 16, L["cout"], fold=(4, 4, self.ncond))
        self._wgrad_join()

    def _vp_numel(self):
        return int(self.vp.theta.numel())

    def optimizer_step(self):
        self._wgrad_join()
        if self.peer is not None and self.peer.grad_peers is not None:
            ops.adam_keras_allreduce_step(self.vp.theta, self.vp.grad, self.vp.adam_m, self.vp.adam_v, self.vp.adam_state,
                                          self.lr, self.peer.grad_peers, self.world, self.dist.rank, self.peer.epoch)
            return
        if self.world > 1:
            self.dist.all_reduce_sum(self.vp.grad)
        ops.adam_keras_step(self.vp.theta, self.vp.grad, self.vp.adam_m, self.vp.adam_v, self.vp.adam_state, self.lr)

    # ------------------------------------------------------------------------------------------
    # whole steps
    # ------------------------------------------------------------------------------------------
    def set_inputs(self, M, cond, eps=None):
        self.M.copy_(M, non_blocking=True)
        self.cond.copy_(cond, non_blocking=True)
        if eps is not None:
            self.eps.copy_(eps, non_blocking=True)
        else:
            self.eps.normal_()

    def _train_body(self):
        if self.peer is not None:
            self.peer.tick()
        if self.overlap_pm:
            # the weight re-pack (latency-bound index shuffling of 1.7 M values) runs beside the input pack (HBM-bound) on the
            # filter-gradient stream; pm(x) needs only the input pack (the perceptual weights are frozen and packed once)
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            if self._wg_side is None:
                self._wg_side = torch.cuda.Stream()
            self._wg_side.wait_stream(main)
            with torch.cuda.stream(self._wg_side):
                self.pack_weights()
            self.pack_inputs()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                if not self._exp_skip_pm0:  # TIMING EXPERIMENT ONLY (ICSG3D_EXP_SKIP_PM0=1)
                    self.pm_forward(0, True, ctx=self.ctx2)
            main.wait_stream(self._wg_side)
            self.encode(True)
            self.decode(True)
            self.pm_forward(1, True)
            main.wait_stream(self._side)
        else:
            self.pack_weights()
            self.pack_inputs()
            self.encode(True)
            self.decode(True)
            self.pm_forward(0, True)
            self.pm_forward(1, True)
        self._defer_taps = self.fuse_tap_loss and not self.fuse_bn_bwd
        self.losses(defer_taps=self._defer_taps)
        self.backward()
        self.optimizer_step()

    def train_step(self):
        """train_on_batch (lattice_vae.py:296-298) on the static input buffers; returns the device metrics tensor."""
        if self.use_graph:
            if self._graph is None:
                self._train_body()  # warm-up: attribute setup, lazy allocations
                torch.cuda.synchronize()
                raise RuntimeError("call capture_train_graph() before train_step() with use_graph")
            if self._segments is not None:
                for kind, obj in self._segments:
                    if kind == "g":
                        obj.replay()
                    else:
                        self.dist.dist.all_reduce(obj, op=self.dist.dist.ReduceOp.SUM, group=self.dist.group)
            else:
                self._graph.replay()
        else:
            self._train_body()
        return self.metrics

    def capture_train_graph(self, snapshot=True):
        """Capture one full train step in a CUDA graph.  The warm-up + capture executes optimiser steps, so the
        parameter/optimiser state is snapshotted and restored around it."""
        saved = None
        if snapshot:
            saved = [t.clone() for t in (self.vp.theta, self.vp.state, self.vp.adam_m, self.vp.adam_v, self.vp.adam_state)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._train_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._segments = None
        if self.world > 1:
            # Data parallel: NCCL stays OUTSIDE the graphs.  Every all-reduce of the step (BatchNorm statistic sums,
            # the flat gradient) ends one captured segment and starts the next; a step = replaying ~50 small graphs
            # with the eager collectives in between (CPU cost ~1 ms/step, hidden behind the GPU).
            segs, pool, cap = [], torch.cuda.graph_pool_handle(), torch.cuda.Stream()
            cur = {}

            def begin():
                cur["g"] = torch.cuda.CUDAGraph()
                cur["ctx"] = torch.cuda.graph(cur["g"], pool=pool, stream=cap)
                cur["ctx"].__enter__()

            def end():
                cur["ctx"].__exit__(None, None, None)
                segs.append(("g", cur["g"]))

            # the eager collective issued between two segments also keeps the ranks in lock step during capture
            dist_obj = self.dist
            real = dist_obj.dist.all_reduce

            def hooked_all_reduce_sum(t):
                end()
                segs.append(("ar", t))
                real(t, op=dist_obj.dist.ReduceOp.SUM, group=dist_obj.group)
                begin()

            orig = dist_obj.all_reduce_sum
            dist_obj.all_reduce_sum = hooked_all_reduce_sum
            try:
                begin()
                self._train_body()
                end()
            finally:
                dist_obj.all_reduce_sum = orig
            g = segs
            self._segments = segs
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._train_body()
        torch.cuda.synchronize()
        if saved is not None:
            for t, sv in zip((self.vp.theta, self.vp.state, self.vp.adam_m, self.vp.adam_v, self.vp.adam_state), saved):
                t.copy_(sv)
        self._graph = g
        self.use_graph = True

    def eval_step(self):
        """test_on_batch (lattice_vae.py:305-310): learning phase 0 -> moving-statistics BN everywhere."""
        self.pack_weights()
        self.pack_inputs()
        self.encode(False)
        self.decode(False)
        self.pm_forward(0, False)
        self.pm_forward(1, False)
        self.losses()
        return self.metrics

    def metrics_host(self):
        m = self.metrics.clone()
        if self.world > 1:
            self.dist.all_reduce_sum(m)
            m /= self.world
        return m.cpu().tolist()
