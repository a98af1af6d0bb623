"""Step executor of the 3-D segmentation U-Net (unet/unet.py:272-355): forward, dual-head losses
(weighted CCE + BCE, unet.py:196-221, 252-259), backward, Keras-Adam, and the learning-phase-0 inference path
(generate.py:220-225).  Like engine.py it only orders C-ABI launches over pre-allocated buffers.

Skip connections never materialise a Concatenate: the BatchNorm-apply pass of the producer writes straight into
its channel slice of the concat buffer (and the UpSampling3D of the other branch into the other slice); in the
backward pass the slices of the concat gradient are read in place.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .engine import BF16, F32, F64, Dist, _BN, _Ctx
from .ops import ACT_NONE, ACT_RELU, POST_NONE, POST_POOL2, POST_UP2, pad16
from .params import ParamStore, unet_specs

# name, cin, cout, level (spatial = d >> level), input source, output destination
#   src:  "x" | name of a block whose BN output feeds it | "pool:<blk>" | "cat:<buf>"
#   dst:  "plain" | ("cat", buf, offset)  [+ pooled copy]  | ("up", buf, offset)
UNET_PLAN = [
    dict(n="c1", cin=4, cout=32, lvl=0, src="x"),
    dict(n="c2", cin=32, cout=64, lvl=0, src="c1", cat=("cat4", 0), pool=True),
    dict(n="c3", cin=64, cout=64, lvl=1, src="pool:c2"),
    dict(n="c4", cin=64, cout=128, lvl=1, src="c3", cat=("cat3", 0), pool=True),
    dict(n="c5", cin=128, cout=128, lvl=2, src="pool:c4"),
    dict(n="c6", cin=128, cout=256, lvl=2, src="c5", cat=("cat2", 0), pool=True),
    dict(n="c9", cin=256, cout=512, lvl=3, src="pool:c6"),
    dict(n="c10", cin=512, cout=512, lvl=3, src="c9", up=("cat2", 256)),
    dict(n="c13", cin=768, cout=512, lvl=2, src="cat:cat2"),
    dict(n="c14", cin=512, cout=256, lvl=2, src="c13", up=("cat3", 128)),
    dict(n="c15", cin=384, cout=256, lvl=1, src="cat:cat3"),
    dict(n="c16", cin=256, cout=128, lvl=1, src="c15", up=("cat4", 64)),
    dict(n="c17", cin=192, cout=128, lvl=0, src="cat:cat4"),
    dict(n="c18", cin=128, cout=128, lvl=0, src="c17"),
]
CAT_CH = {"cat2": (2, 768), "cat3": (1, 384), "cat4": (0, 192)}


class UNetEngine:
    def __init__(self, batch, d=32, channels=4, classes=95, device="cuda", params: ParamStore | None = None, lr=1e-6,
                 class_weight=None, dist: Dist | None = None, seed=2, train=True):
        assert channels in (1, 4), "the reference feeds 4 channels (density + coordinates) or density only"
        assert d % 8 == 0 and d & (d - 1) == 0
        self.B, self.d, self.channels, self.classes = batch, d, channels, classes
        self.dev = torch.device(device)
        self.lr = float(lr)
        self.dist = dist
        self.world = dist.world if dist else 1
        dev = self.dev
        self.pp = params or ParamStore(unet_specs(channels, classes), dev, with_grads=train, with_adam=train).init(seed)
        # split-K scratch of the per-tap conv kernel (the 4^3 / 8^3 layers have too few output tiles to fill 148 SMs
        # otherwise): largest request over every layer's fprop and dgrad operand shape
        ws_need = 0
        for spec in UNET_PLAN:
            cin_eff = pad16(channels if spec["n"] == "c1" else spec["cin"])
            D_ = d >> spec["lvl"]
            ws_need = max(ws_need, ops.conv3d_k3_workspace_bytes(batch, D_, cin_eff, spec["cout"]),
                          ops.conv3d_k3_workspace_bytes(batch, D_, spec["cout"], cin_eff))
            if spec["src"].startswith("cat:"):  # data gradient of the skip half alone (Upsample-into-Conv fold)
                cs = next(x["cout"] for x in UNET_PLAN if x.get("cat", (None,))[0] == spec["src"][4:])
                ws_need = max(ws_need, ops.conv3d_k3_workspace_bytes(batch, D_, spec["cout"], cs))
        self.ctx = _Ctx(dev, conv_ws_bytes=ws_need)
        self.train_enabled = train
        B = batch
        z = lambda *s, dt=BF16: torch.zeros(*s, dtype=dt, device=dev)
        self.X = z(B, d, d, d, channels, dt=F32)
        self.species = z(B, d, d, d, dt=torch.uint8)
        self.x16 = z(B, d, d, d, 16)
        self.cat = {k: z(B, d >> lvl, d >> lvl, d >> lvl, c) for k, (lvl, c) in CAT_CH.items()}
        self.dcat = {k: z(B, d >> lvl, d >> lvl, d >> lvl, c) for k, (lvl, c) in CAT_CH.items()} if train else {}
        self.L = {}
        for spec in UNET_PLAN:
            n, cin, cout, lvl = spec["n"], spec["cin"], spec["cout"], spec["lvl"]
            D = d >> lvl
            cin_eff = channels if n == "c1" else cin
            L = dict(spec)
            L.update(D=D, cin_pad=pad16(cin_eff), cin_real=cin_eff, a=z(B, D, D, D, cout), bn=_BN(cout, dev),
                     wf=z(27, cout, pad16(cin_eff)))
            if "cat" in spec:
                buf, off = spec["cat"]
                L["y"] = self.cat[buf][..., off:off + cout]
            elif "up" in spec:
                buf, off = spec["up"]
                L["y"] = self.cat[buf][..., off:off + cout]  # written upsampled x2
            else:
                L["y"] = z(B, D, D, D, cout)
            if spec.get("pool"):
                L["p"] = z(B, D // 2, D // 2, D // 2, cout)
                L["idx"] = z(B, D // 2, D // 2, D // 2, cout, dt=torch.uint8)
            if train:
                L["dc"] = z(B, D, D, D, cout)
                if n != "c1":
                    L["wd"] = z(27, pad16(cin_eff), cout)
                # gradient w.r.t. this block's (un-pooled, un-upsampled-domain) output when it is a plain tensor
                if "cat" not in spec and "up" not in spec:
                    L["dy"] = z(B, D, D, D, cout)
                if spec.get("pool"):
                    L["dp"] = z(B, D // 2, D // 2, D // 2, cout)
            self.L[n] = L
        # Upsample-into-Conv fold for learning phase 0 (csrc/conv3d_upfold.cu): consumer block -> (skip slice, low tensor)
        self.fold = os.environ.get("ICSG3D_UPFOLD", "1") != "0"
        # ... and in the train step for the forward and the data gradient (the filter gradient still reads the
        # materialised upsampled tensor, so the BatchNorm pass of the producer writes both forms)
        self.fold_train = train and self.fold and os.environ.get("ICSG3D_UPFOLD_TRAIN", "1") != "0"
        self._wfold_fresh = False
        for L in self.L.values():
            if "up" in L:
                L["ylow"] = z(B, L["D"], L["D"], L["D"], L["cout"])
                L["dylow"] = z(B, L["D"], L["D"], L["D"], L["cout"]) if self.fold_train else None
        for L in self.L.values():
            if L["src"].startswith("cat:"):
                buf = L["src"][4:]
                sk = next(x for x in self.L.values() if x.get("cat", (None,))[0] == buf)
                up = next(x for x in self.L.values() if x.get("up", (None,))[0] == buf)
                L["fold"] = (sk["y"], up["ylow"])
                L["fold_ch"] = (sk["cat"][1], sk["cout"], up["up"][1], up["cout"])
                L["fold_up"] = up["n"]
                # in the TRAIN step the fold pays where the low-resolution grid alone fills the SMs (>= 128 tiles of 128
                # voxels): the folded weights are re-packed every step and the deep layers' data gradient has too few
                # tiles without a K split (measured at batch 8: folding c13 / c15 too costs 0.4 ms, c17 alone gains)
                L["fold_tr"] = up["fold_tr"] = self.fold_train and B * (L["D"] // 2) ** 3 >= 16384
                s0, cs, u0, cu = L["fold_ch"]
                lib = ops._lib.lib()
                L["wfold"] = z(int(lib.icsg3d_conv3d_upfold_wpack_elems(cs, cu, L["cout"])))
                L["wd_skip"] = z(27, cs, L["cout"]) if L["fold_tr"] else None
                L["wdlow"] = z(int(lib.icsg3d_conv3d_upfold_dgrad_wpack_elems(L["cout"], cu))) if L["fold_tr"] else None
        cmax = max(spec["cout"] for spec in UNET_PLAN)
        self.one, self.zero = torch.ones(cmax, dtype=F32, device=dev), torch.zeros(cmax, dtype=F32, device=dev)
        # heads
        self.nout_h = pad16(classes + 1)
        self.h_wf, self.h_wd = z(1, self.nout_h, 128), z(1, 128, self.nout_h)
        self.h_bias = z(self.nout_h, dt=F32)
        self.logits = z(B, d, d, d, self.nout_h, dt=F32)
        self.argmax = z(B, d, d, d, dt=torch.uint8)
        self.mask = z(B, d, d, d, dt=torch.uint8)
        self.sigp = z(B, d, d, d, dt=F32)
        self.h_nparts = ops.heads_loss_nparts(B * d ** 3)
        self.h_partials = torch.zeros(self.h_nparts, 6, dtype=F64, device=dev)
        self.h_partials_fused = torch.zeros(ops.heads_loss_fused_nparts(B * d ** 3), 6, dtype=F64, device=dev)
        self.fuse_heads = os.environ.get("ICSG3D_FUSE_HEADS", "1") != "0"
        # filter / bias gradients on a side stream — single process only: the data-parallel step interleaves eager NCCL
        # all-reduces with its launches, and a 2-rank run of the bench with the side stream on did not finish inside its
        # time limit (the round's last GPU call; not diagnosed), so under NCCL everything stays on one stream, as in the
        # VAE engine's NCCL mode
        self.overlap_wgrad = self.world == 1 and os.environ.get("ICSG3D_UNET_OVERLAP_WGRAD", "1") != "0"
        self._wg_side, self._wg_pending = None, False
        self.keep_logits = False  # diagnostics / parity tests: also materialise the fp32 head logits the fused kernel skips
        self.h_raw = torch.zeros(6, dtype=F64, device=dev)
        self.metrics = torch.zeros(5, dtype=F32, device=dev)
        cw = torch.full((classes,), float(classes)) if class_weight is None else torch.as_tensor(class_weight, dtype=F32)
        # reference quirk (unet.py:254): compile() passes weighted_categorical_crossentropy(num_classes), i.e. the
        # SCALAR 95.0 as weight for every class; a (classes,) vector reproduces custom_objects["loss"] instead.
        self.class_w = cw.to(dev).float().contiguous()
        if train:
            self.dlogits = z(B, d, d, d, self.nout_h)
            self.dwcat = z(1, 128, self.nout_h, dt=F32)
            self.colsum = torch.zeros(2 * 512, dtype=F64, device=dev)
            self.side_colsum = torch.zeros(2 * 512, dtype=F64, device=dev)
            self.side_partials = torch.zeros_like(self.ctx.partials)
        self._graph = None
        self.use_graph = False
        self.wg_ws = None
        if train:  # filter-gradient scratch owned by this engine (see VAEEngine.wg_ws)
            need = max(ops.conv3d_k3_wgrad_workspace_bytes(B, L["D"], L["cin_pad"], L["cout"]) for L in self.L.values())
            need = max(need, ops.conv3d_k3_wgrad_workspace_bytes(B, d, 128, self.nout_h, k1=True))
            self.wg_ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------------------------------
    def pack_weights(self, dgrad=True):
        p = self.pp.p
        self._wfold_fresh = not dgrad and self.fold
        for n, L in self.L.items():
            ops.pack_conv_w_fprop(p[n + "/kernel"], cin_pad=L["cin_pad"], out=L["wf"])
            if "fold" in L and ((not dgrad and self.fold) or (dgrad and L["fold_tr"])):
                # folded taps of the upsampled channels (+ in training the data-gradient operands of both halves)
                L["wfold"] = ops.pack_conv_w_upfold(p[n + "/kernel"], *L["fold_ch"], out=L["wfold"])
                if dgrad:
                    s0, cs, u0, cu = L["fold_ch"]
                    L["wd_skip"] = ops.pack_conv_w_dgrad_slice(p[n + "/kernel"], s0, cs, out=L["wd_skip"])
                    L["wdlow"] = ops.pack_conv_w_upfold_dgrad(p[n + "/kernel"], u0, cu, out=L["wdlow"])
                    continue
            if dgrad and "wd" in L:
                ops.pack_conv_w_dgrad(p[n + "/kernel"], cin_pad=L["cin_pad"], out=L["wd"])
        ops.pack_heads_w(p["soft/kernel"], p["sig/kernel"], p["soft/bias"], p["sig/bias"], self.h_wf, self.h_wd, self.h_bias)

    def _input_of(self, L):
        src = L["src"]
        if src == "x":
            return self.x16
        if src.startswith("pool:"):
            return self.L[src[5:]]["p"]
        if src.startswith("cat:"):
            return self.cat[src[4:]]
        return self.L[src]["y"]

    def _inference_coeffs(self, L):
        p, n, st = self.pp.p, L["n"], L["bn"]
        ops.bn_inference_coeffs(p[f"bn_{n}/gamma"], p[f"bn_{n}/beta"], p[f"bn_{n}/moving_mean"], p[f"bn_{n}/moving_variance"],
                                st.scale, st.shift)

    def inference_coeffs(self):
        """BatchNorm scale / shift of every block from the moving statistics (learning phase 0); they change only with
        the weights, so a serving loop computes them once (predict(repack=False) / forward(coeffs=False))."""
        for L in self.L.values():
            self._inference_coeffs(L)

    def _bn_fwd(self, L, training, part=None, coeffs=True):
        p = self.pp.p
        n, C, st, x = L["n"], L["cout"], L["bn"], L["a"]
        g, b = p[f"bn_{n}/gamma"], p[f"bn_{n}/beta"]
        mm, mv = p[f"bn_{n}/moving_mean"], p[f"bn_{n}/moving_variance"]
        if training:
            rows = x.numel() // C
            if part is None:  # no fused statistics from the conv epilogue for this layer shape
                np_ = ops.bn_nparts(rows, C, x.dtype)
                part = self.ctx.partials[: np_ * 2 * C].view(np_, 2, C)
                ops.bn_stats(x, C, part)
            if self.world > 1:
                ops.bn_reduce_partials(part, st.sums)
                self.dist.all_reduce_sum(st.sums)
                ops.bn_finalize(st.sums, float(rows * self.world), g, b, st.mean, st.rstd, st.scale, st.shift, mm, mv)
            else:
                ops.bn_reduce_finalize(part, float(rows), g, b, st.sums, st.mean, st.rstd, st.scale, st.shift, mm, mv)
        elif coeffs:
            ops.bn_inference_coeffs(g, b, mm, mv, st.scale, st.shift)
        if "up" in L:
            ops.bn_apply_fwd(x, C, st.scale, st.shift, ACT_NONE, POST_UP2, y=L["y"])
            if training and L.get("fold_tr"):  # the folded consumer reads the low-resolution output
                ops.bn_apply_fwd(x, C, st.scale, st.shift, ACT_NONE, POST_NONE, y=L["ylow"])
        else:
            ops.bn_apply_fwd(x, C, st.scale, st.shift, ACT_NONE, POST_NONE, y=L["y"])
            if L.get("pool"):
                ops.bn_apply_fwd(x, C, st.scale, st.shift, ACT_NONE, POST_POOL2, y=L["p"], pool_idx=L["idx"])

    def forward(self, training, want_probs=None, with_grad=False, x_packed=False, losses=True, heads=True, coeffs=True):
        """unet_3d_multiclass (unet.py:272-355) on self.X [+ losses on self.species].  x_packed: self.x16 already holds
        the bf16 input (written by the producer, e.g. utils.lattice_params_device's fused pack); losses=False stops at
        the head logits (inference: generate.py:220); coeffs=False (learning phase 0): the BatchNorm scale / shift of the
        moving statistics are already in place (inference_coeffs() after the last weight change)."""
        p = self.pp.p
        if x_packed:
            pass
        elif self.channels == 4:
            ops.pack_vae_input(self.X, None, None, self.x16)
        else:
            ops.f32_to_bf16_rows(self.X, 1, self.x16)
        fold = self.fold and self._wfold_fresh  # folded weights are packed by pack_weights(dgrad=False) only
        for L in self.L.values():
            xin, part = self._input_of(L), None
            if not training and ("up" not in L or fold):
                # learning phase 0: BatchNorm is a fixed per-channel affine -> applied in the conv epilogue, which writes
                # the block output (a plain tensor or its slice of a concatenation buffer) directly.  With the
                # Upsample-into-Conv fold the blocks before an UpSampling3D keep their LOW-resolution output
                # and the consumer convolves it with 8 folded taps per output phase: no upsampled tensor, no BN pass.
                st = L["bn"]
                if coeffs:
                    self._inference_coeffs(L)
                dst = L["ylow"] if "up" in L else L["y"]
                if "fold" in L and fold:
                    skip, low = L["fold"]
                    ops.conv3d_k3_upfold(skip, low, L["wfold"], p[L["n"] + "/bias"], L["cout"], act=ACT_RELU, out=dst,
                                         post=(st.scale, st.shift), tag=f"unet.{L['n']}.fprop",
                                         nominal=(L["cin_real"], L["cout"]))
                else:
                    ops.conv3d_k3(xin, L["wf"], p[L["n"] + "/bias"], out=dst, act=ACT_RELU, post=(st.scale, st.shift),
                                  ws=self.ctx.conv_ws, tag=f"unet.{L['n']}.fprop", nominal=(L["cin_real"], L["cout"]))
                if L.get("pool"):
                    ops.bn_apply_fwd(L["y"], L["cout"], self.one, self.zero, ACT_NONE, POST_POOL2, y=L["p"])
                continue
            if training and L.get("fold_tr") and "fold" in L:
                # train step: the upsampled half of the input as 8 folded taps on the producer's low-resolution output
                skip, low = L["fold"]
                ops.conv3d_k3_upfold(skip, low, L["wfold"], p[L["n"] + "/bias"], L["cout"], act=ACT_RELU, out=L["a"],
                                     tag=f"unet.{L['n']}.fprop", nominal=(L["cin_real"], L["cout"]))
                self._bn_fwd(L, training, part=None, coeffs=coeffs)
                continue
            if training:  # BatchNorm statistics out of the conv epilogue where the serving kernel has them
                nparts = ops.conv3d_k3_stats_parts(xin, L["wf"])
                if nparts > 0:
                    part = self.ctx.partials[: nparts * 2 * L["cout"]].view(nparts, 2, L["cout"])
            ops.conv3d_k3(xin, L["wf"], p[L["n"] + "/bias"], out=L["a"], act=ACT_RELU, stats=part,
                          ws=self.ctx.conv_ws, tag=f"unet.{L['n']}.fprop", nominal=(L["cin_real"], L["cout"]))
            self._bn_fwd(L, training, part=part, coeffs=coeffs)
        if not heads:
            return
        M = self.B * self.d ** 3
        if losses and want_probs is None and self.fuse_heads:
            # head GEMM + losses + metric counts + bf16 gradient in one kernel: the fp32 logits never exist in HBM
            ops.heads_loss_fused(self.L["c18"]["y"], self.h_wf, self.h_bias, self.classes, self.species, self.class_w,
                                 1.0 / (M * self.world), self.h_partials_fused, argmax_out=self.argmax, sig_prob=self.sigp,
                                 dlogits=self.dlogits if with_grad else None)
            ops.heads_loss_finalize(self.h_partials_fused, float(M), self.metrics, self.h_raw)
            if self.keep_logits:
                ops.conv3d_k3(self.L["c18"]["y"], self.h_wf, self.h_bias, out=self.logits)
            return
        ops.conv3d_k3(self.L["c18"]["y"], self.h_wf, self.h_bias, out=self.logits, tag="unet.heads.fprop",
                      nominal=(128, self.classes + 1))
        if not losses:
            return
        M = self.B * self.d ** 3
        ops.heads_loss(self.logits, self.classes, self.species, self.class_w, 1.0 / (M * self.world), self.h_partials,
                       argmax_out=self.argmax, sig_prob=self.sigp, dlogits=self.dlogits if with_grad else None,
                       probs=want_probs)
        ops.heads_loss_finalize(self.h_partials, float(M), self.metrics, self.h_raw)

    # ------------------------------------------------------------------------------------------
    def _bias_grad(self, dc, C, gbias):
        """db = column sums of dc.  Only Adam needs it: with the filter gradients on the side stream, on scratch of its own
        (the main stream's BatchNorm kernels keep using ctx.partials / colsum meanwhile)."""
        rows = dc.numel() // C
        n = ops.bn_nparts(rows, C, dc.dtype)
        with self._wgrad_stream():
            side = self.overlap_wgrad
            part = (self.side_partials if side else self.ctx.partials)[: n * 2 * C].view(n, 2, C)
            ops.bn_stats(dc, C, part)
            ops.bn_reduce_grads(part, (self.side_colsum if side else self.colsum)[: 2 * C], None, gbias)

    def _grad_wrt_output(self, L):
        """(dy, post, idx, dy2) describing the gradient reaching block L's BN output."""
        if "up" in L:
            if L.get("fold_tr"):  # written at low resolution by the folded consumer's data-gradient kernel
                return L["dylow"], POST_NONE, None, None
            buf, off = L["up"]
            return self.dcat[buf][..., off:off + L["cout"]], POST_UP2, None, None
        if "cat" in L:
            buf, off = L["cat"]
            return L["dp"], POST_POOL2, L["idx"], self.dcat[buf][..., off:off + L["cout"]]
        return L["dy"], POST_NONE, None, None

    def _dst_of_input_grad(self, L):
        src = L["src"]
        if src.startswith("pool:"):
            return self.L[src[5:]]["dp"]
        if src.startswith("cat:"):
            return self.dcat[src[4:]]
        return self.L[src]["dy"]

    def backward(self):
        p, g = self.pp.p, self.pp.g
        c18 = self.L["c18"]
        # heads: dW, db, and the gradient w.r.t. c18's BN output
        ops.conv3d_k1_wgrad(c18["y"], self.dlogits, cin=128, cout=self.nout_h, out=self.dwcat, ws=self.wg_ws)
        rows = self.dlogits.numel() // self.nout_h
        n = ops.bn_nparts(rows, self.nout_h, BF16)
        part = self.ctx.partials[: n * 2 * self.nout_h].view(n, 2, self.nout_h)
        ops.bn_stats(self.dlogits, self.nout_h, part)
        ops.bn_reduce_partials(part, self.colsum[: 2 * self.nout_h])
        ops.unpack_heads_grad(self.dwcat, self.colsum, self.classes, g["soft/kernel"], g["sig/kernel"], g["soft/bias"],
                              g["sig/bias"])
        ops.conv3d_k3(self.dlogits, self.h_wd, None, out=c18["dy"], tag="unet.heads.dgrad", nominal=(self.classes + 1, 128))
        for L in reversed(list(self.L.values())):
            nme, C, st = L["n"], L["cout"], L["bn"]
            dy, post, idx, dy2 = self._grad_wrt_output(L)
            x = L["a"]
            nb = ops.bn_bwd_nparts(x, C, post)
            part = self.ctx.partials[: nb * 2 * C].view(nb, 2, C)
            ops.bn_bwd_reduce(dy, x, C, st.mean, st.rstd, st.scale, st.shift, ACT_NONE, post, idx, part, dy2=dy2)
            ops.bn_reduce_grads(part, st.bsums, g[f"bn_{nme}/gamma"], g[f"bn_{nme}/beta"])
            sums = st.bsums
            if self.world > 1:
                st.bsums_g.copy_(st.bsums)
                self.dist.all_reduce_sum(st.bsums_g)
                sums = st.bsums_g
            rws = x.numel() // C
            ops.bn_bwd_apply(dy, x, C, st.mean, st.rstd, st.scale, st.shift, ACT_NONE, post, idx, sums, float(rws * self.world),
                             L["dc"], pre_relu=True, dy2=dy2)
            self._bias_grad(L["dc"], C, g[nme + "/bias"])
            xin = self._input_of(L)
            gk = g[nme + "/kernel"]
            # filter gradient: only Adam needs it -> on a side stream under the BatchNorm-backward / data-gradient chain of
            # the next layers (HBM-bound passes next to an L2-bound tensor kernel; the per-tap kernels leave registers
            # and shared memory for a BatchNorm block on the same SM)
            with self._wgrad_stream():
                if L["cin_real"] == L["cin_pad"]:
                    ops.conv3d_k3_wgrad(xin, L["dc"], cin=L["cin_pad"], cout=C, out=gk.view(27, L["cin_real"], C),
                                        tag=f"unet.{nme}.wgrad", ws=self.wg_ws)
                else:
                    scratch = self.ctx.dw_pad[: 27 * L["cin_pad"] * C].view(27, L["cin_pad"], C)
                    ops.conv3d_k3_wgrad(xin, L["dc"], cin=L["cin_pad"], cout=C, out=scratch, tag=f"unet.{nme}.wgrad",
                                        nominal=(L["cin_real"], C), ws=self.wg_ws)
                    ops.unpack_conv_dw(scratch, L["cin_real"], C, out=gk)
            if L.get("fold_tr") and "fold" in L:
                # data gradient of the folded layer: 27 taps for the skip channels only, 64 folded (phase, tap) pairs
                # straight to the LOW-resolution gradient of the upsampled producer
                s0, cs, u0, cu = L["fold_ch"]
                ops.conv3d_k3(L["dc"], L["wd_skip"], None, out=self._dst_of_input_grad(L)[..., s0:s0 + cs], ws=self.ctx.conv_ws,
                              tag=f"unet.{nme}.dgrad", nominal=(C, cs))
                ops.conv3d_k3_upfold_dgrad_low(L["dc"], L["wdlow"], cu, out=self.L[L["fold_up"]]["dylow"],
                                               tag=f"unet.{nme}.dgrad_low")
            elif nme != "c1":
                ops.conv3d_k3(L["dc"], L["wd"], None, out=self._dst_of_input_grad(L), ws=self.ctx.conv_ws,
                              tag=f"unet.{nme}.dgrad")
        if self._wg_pending:
            torch.cuda.current_stream().wait_stream(self._wg_side)
            self._wg_pending = False

    def _wgrad_stream(self):
        """Context of the filter-gradient launches: the side stream (forked from the current stream here, joined at the
        end of backward()) or, with overlap_wgrad off, the current stream."""
        import contextlib
        if not self.overlap_wgrad:
            return contextlib.nullcontext()
        if self._wg_side is None:
            self._wg_side = torch.cuda.Stream()
        self._wg_side.wait_stream(torch.cuda.current_stream())
        self._wg_pending = True
        return torch.cuda.stream(self._wg_side)

    def optimizer_step(self):
        if self.world > 1:
            self.dist.all_reduce_sum(self.pp.grad)
        ops.adam_keras_step(self.pp.theta, self.pp.grad, self.pp.adam_m, self.pp.adam_v, self.pp.adam_state, self.lr)

    # ------------------------------------------------------------------------------------------
    def set_inputs(self, X, species=None):
        self.X.copy_(X.reshape(self.X.shape), non_blocking=True)
        if species is not None:
            self.species.copy_(species.reshape(self.species.shape), non_blocking=True)

    def _train_body(self):
        self.pack_weights()
        self.forward(True, with_grad=True)
        self.backward()
        self.optimizer_step()

    def train_step(self):
        """train_on_batch(X, [onehot(S), S != 0]) -> metrics [loss, soft, sig, f1_m, wr_m] (device tensor)."""
        self._wfold_fresh = False  # the weights change: the inference-time folded pack is stale (a graph replay skips the
        if self.use_graph and self._graph is not None:  # Python side of pack_weights)
            self._graph.replay()
        else:
            self._train_body()
        return self.metrics

    def capture_train_graph(self):
        if self.world > 1:
            # the data-parallel U-Net step issues NCCL all-reduces (BatchNorm sums, flat gradient); capturing those inside
            # one graph is not covered by a test, so the data-parallel U-Net runs eagerly
            self.use_graph = False
            return
        saved = [t.clone() for t in (self.pp.theta, self.pp.state, self.pp.adam_m, self.pp.adam_v, self.pp.adam_state)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._train_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._train_body()
        torch.cuda.synchronize()
        for t, sv in zip((self.pp.theta, self.pp.state, self.pp.adam_m, self.pp.adam_v, self.pp.adam_state), saved):
            t.copy_(sv)
        self._graph, self.use_graph = g, True

    def eval_step(self):
        self.pack_weights(dgrad=False)
        self.forward(False)
        return self.metrics

    def predict(self, probs_out=None, x_packed=False, threshold=0.8, repack=True):
        """Learning phase 0 forward (generate.py:220-225): fills self.argmax (species labels), self.mask
        (sigmoid >= threshold) and self.sigp; with probs_out also the softmax probabilities `model.predict` returns."""
        if repack:
            self.pack_weights(dgrad=False)
        if probs_out is None:  # labels only: head GEMM + soft-max arg-max + threshold in one kernel, no logits in HBM
            self.forward(False, x_packed=x_packed, losses=False, heads=False, coeffs=repack)
            ops.heads_predict_fused(self.L["c18"]["y"], self.h_wf, self.h_bias, self.classes, float(threshold),
                                    argmax=self.argmax, mask=self.mask, sig_prob=self.sigp)
            return
        self.forward(False, want_probs=probs_out, x_packed=x_packed, losses=True, coeffs=repack)
        ops.heads_predict(self.logits, self.classes, float(threshold), argmax=self.argmax, mask=self.mask, sig_prob=self.sigp)

    def metrics_host(self):
        m = self.metrics.clone()
        if self.world > 1:
            self.dist.all_reduce_sum(m)
            m /= self.world
        return m.cpu().tolist()
