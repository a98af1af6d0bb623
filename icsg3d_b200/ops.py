"""Torch-tensor front-end of the C ABI (include/icsg3d.h).

PyTorch is used here only as the device allocator and stream provider: every function validates its
tensors, extracts raw device pointers + the current CUDA stream and calls into libicsg3d.so.
Nothing in this module computes on the CPU and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
POST_NONE, POST_POOL2, POST_UP2 = 0, 1, 2
DT_BF16, DT_F32 = 0, 1
LEAKY_ALPHA = 0.3  # Keras LeakyReLU() default


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.Icsg3dError("icsg3d ops need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ld(t):
    """Channel stride (elements) of one voxel: supports channel slices of a wider NDHWC tensor (concat buffers)."""
    return t.stride(-2) if t.dim() >= 2 else t.shape[-1]


def _chk(t, dtype, name):
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name}: channels must be contiguous")
    if t.dim() == 5:
        B, D, H, W, _ = t.shape
        ld = t.stride(-2)
        if t.stride(2) != W * ld or t.stride(1) != H * W * ld or t.stride(0) != D * H * W * ld:
            raise ValueError(f"{name}: only channel-sliced views of a contiguous NDHWC tensor are supported")
    elif not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")


def launch_count() -> int:
    return int(_lib.lib().icsg3d_launch_count())


def pad16(c: int) -> int:
    return (c + 15) // 16 * 16


# ------------------------------------------------------------------------------------------------
# Conv3D
# ------------------------------------------------------------------------------------------------
def pack_conv_w_fprop(w, cin_pad=None, cout_pad=None, cin_lead=0, fold=1, fold_c=0, out=None):
    """w: fp32 (3,3,3,Cin,Cout) Keras layout -> bf16 [27][cout_pad][cin_pad]."""
    _chk(w, torch.float32, "w")
    cin, cout = w.shape[3], w.shape[4]
    if cin_pad is None:
        cin_pad = pad16(cin if fold <= 1 else cin_lead + fold_c)
    if cout_pad is None:
        cout_pad = pad16(cout)
    if out is None:
        out = torch.empty((27, cout_pad, cin_pad), dtype=torch.bfloat16, device=w.device)
    _lib.call("icsg3d_pack_conv_w_fprop", _ptr(w), _ptr(out), cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c,
              _stream())
    return out


def pack_conv_w_dgrad(w, cin_pad=None, cout_pad=None, out=None):
    """w: fp32 (3,3,3,Cin,Cout) -> bf16 [27][cin_pad][cout_pad] with mirrored taps."""
    _chk(w, torch.float32, "w")
    cin, cout = w.shape[3], w.shape[4]
    cin_pad = cin_pad or pad16(cin)
    cout_pad = cout_pad or pad16(cout)
    if out is None:
        out = torch.empty((27, cin_pad, cout_pad), dtype=torch.bfloat16, device=w.device)
    _lib.call("icsg3d_pack_conv_w_dgrad", _ptr(w), _ptr(out), cin, cout, cin_pad, cout_pad, _stream())
    return out


def pack_jobs_table(jobs, device):
    """jobs: list of (w fp32 (3,3,3,Cin,Cout), wpack bf16, mode 0|1, cin_lead, fold, fold_c) -> device int64 table."""
    rows = []
    for w, wp, mode, cin_lead, fold, fold_c in jobs:
        _chk(w, torch.float32, "w")
        _chk(wp, torch.bfloat16, "wpack")
        cin, cout = w.shape[3], w.shape[4]
        if mode == 0:
            cout_pad, cin_pad = wp.shape[1], wp.shape[2]
        elif mode == 1:
            cout_pad, cin_pad = wp.shape[2], wp.shape[1]
        elif mode == 4:  # lean split pack of the encoder's first conv [27][cout_pad][32]
            cout_pad, cin_pad = wp.shape[1], 16
        elif mode == 2:  # split fprop operand [27][cout_pad][3*cin_pad]
            cout_pad, cin_pad = wp.shape[1], wp.shape[2] // 3
        else:            # split dgrad operand [27][cin_pad][3*cout_pad]
            cout_pad, cin_pad = wp.shape[2] // 3, wp.shape[1]
        rows.append([w.data_ptr(), wp.data_ptr(), cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c, mode])
    return torch.tensor(rows, dtype=torch.int64).to(device)


def pack_conv_w_batch(table, max_blocks=32):
    _lib.call("icsg3d_pack_conv_w_batch", _ptr(table), table.shape[0], max_blocks, _stream())


def unpack_conv_dw(dw_pad, cin, cout, cin_lead=0, fold=1, fold_c=0, out=None):
    _chk(dw_pad, torch.float32, "dw_pad")
    cin_pad, cout_pad = dw_pad.shape[1], dw_pad.shape[2]
    if out is None:
        out = torch.empty((3, 3, 3, cin, cout), dtype=torch.float32, device=dw_pad.device)
    _lib.call("icsg3d_unpack_conv_dw", _ptr(dw_pad), _ptr(out), cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c,
              _stream())
    return out


# Optional per-launch timing of the dominant kernel (bench.py's roofline pass): when TIMING is a list, every
# tcgen05 conv launch appends (tag, algorithmic_flops, start_event, end_event).
TIMING = None


def _timed(tag, flops, executed=None):
    """flops: algorithmic (nominal channels); executed: what the launch really multiplies (padded / folded
    / split channel counts) — reported side by side by bench.py."""
    class _T:
        def __enter__(self):
            if TIMING is not None:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
            return self

        def __exit__(self, *a):
            if TIMING is not None:
                self.e1.record()
                TIMING.append((tag, flops, self.e0, self.e1, flops if executed is None else executed))
    return _T()


def _nbytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def conv3d_k3_workspace_bytes(B, D, cin, nout):
    return int(_lib.lib().icsg3d_conv3d_k3_workspace_bytes(B, D, D, D, cin, nout))


def conv3d_k3_stats_parts(x, wpack):
    """Rows of BatchNorm partials the fused conv+statistics path writes for this layer shape (0 = not available)."""
    B, D, H, W, _ = x.shape
    return int(_lib.lib().icsg3d_conv3d_k3_stats_parts(B, D, H, W, wpack.shape[2], wpack.shape[1]))


def conv3d_k3(x, wpack, bias=None, *, cin=None, n_store=None, act=ACT_NONE, alpha=LEAKY_ALPHA, out=None,
              out_dtype=torch.bfloat16, ref=False, nominal=None, tag="conv", stats=None, ws=None, split=False, fmt=None,
              post=None):
    """x: bf16 [B,D,H,W,ldx]; wpack: bf16 [27][nout][cin]; returns y [B,D,H,W,n_store].

    `ref=True` runs the CUDA-core cross-check kernel (fp32 output) instead of the tcgen05 kernel.
    `stats`: fp64 [parts, 2, nout] with parts = conv3d_k3_stats_parts(x, wpack) > 0 — the kernel also writes the
    per-CTA BatchNorm partials (sum, sum of squares) of the stored output.
    `split=True`: x / wpack are fp32-class split operands (pack_conv_w_fprop_x3 etc.) in the format `fmt` (default
    ops.SPLIT_FMT): fp16 pairs need the conv to be told (operand-format field, output rescale); bf16 pairs (fmt 0) are
    ordinary operands with 3x the channels.
    `ws`: optional scratch tensor (any dtype, one per stream): layers with too few tiles (4^3 / 2^3 grids) split K over it.
    `post=(scale, shift)`: fp32 [nout] per-channel affine applied after the activation in the epilogue (the inference
    BatchNorm that follows Conv3D+ReLU in the U-Net): y = scale * act(conv + bias) + shift.
    """
    _chk(x, torch.bfloat16, "x")
    _chk(wpack, torch.bfloat16, "wpack")
    B, D, H, W, _ = x.shape
    ldx = _ld(x)
    nout, cin_w = wpack.shape[1], wpack.shape[2]
    cin = cin or cin_w
    if cin != cin_w:
        raise ValueError(f"conv3d_k3: cin {cin} does not match packed weights {cin_w}")
    n_store = n_store or nout
    if bias is not None:
        # a bias shorter than the padded nout is fine: it is a view into a flat parameter buffer with slack
        # (params.ParamStore.PAD) and the columns >= n_store it would feed are never stored.
        _chk(bias, torch.float32, "bias")
    if ref:
        y = torch.empty((B, D, H, W, nout), dtype=torch.float32, device=x.device) if out is None else out
        _lib.call("icsg3d_ref_conv3d_k3", _ptr(x), ldx, _ptr(wpack), _ptr(bias), _ptr(y), y.shape[-1], B, D, H, W, cin,
                  nout, act, alpha, _stream())
        return y
    if out is None:
        out = torch.empty((B, D, H, W, n_store), dtype=out_dtype, device=x.device)
    ydt = DT_BF16 if out.dtype == torch.bfloat16 else DT_F32
    nc, no = nominal if nominal else (cin, nout)
    kind = "igemm"
    if TIMING is not None and wpack.shape[0] == 27:
        plan = (ctypes.c_int * 10)()
        _lib.lib().icsg3d_conv3d_k3_plan(B, D, H, W, cin, nout, 148, plan)
        kind = ("pertap", "halo", "stream")[plan[0]]
    with _timed((kind, tag), 2.0 * B * D * H * W * wpack.shape[0] * nc * no, 2.0 * B * D * H * W * wpack.shape[0] * cin * nout):
        if split and (SPLIT_FMT if fmt is None else fmt) == 1:
            fn = "icsg3d_conv3d_k1_igemm_f16" if wpack.shape[0] == 1 else "icsg3d_conv3d_k3_igemm_f16"
            _lib.call(fn, _ptr(x), ldx, _ptr(wpack), _ptr(bias), _ptr(out), _ld(out), ydt, n_store, B, D, H, W, cin, nout,
                      act, alpha, 1.0 / SPLIT_WSCALE, _stream())
        elif post is not None:
            if wpack.shape[0] != 27 or split or stats is not None:
                raise ValueError("conv3d_k3: post= is for plain 3x3x3 layers without fused statistics")
            _chk(post[0], torch.float32, "post scale")
            _chk(post[1], torch.float32, "post shift")
            _lib.call("icsg3d_conv3d_k3_igemm_post", _ptr(x), ldx, _ptr(wpack), _ptr(bias), _ptr(post[0]), _ptr(post[1]),
                      _ptr(out), _ld(out), ydt, n_store, B, D, H, W, cin, nout, act, alpha, _ptr(ws),
                      ctypes.c_int64(ws.numel() * ws.element_size() if ws is not None else 0), _stream())
        elif stats is not None:
            _chk(stats, torch.float64, "stats")
            if stats.dim() != 3 or stats.shape[1] != 2 or stats.shape[2] != nout or not stats.is_contiguous():
                raise ValueError("conv3d_k3: stats must be a contiguous [parts, 2, nout] fp64 tensor")
            _lib.call("icsg3d_conv3d_k3_igemm_stats", _ptr(x), ldx, _ptr(wpack), _ptr(bias), _ptr(out), _ld(out), ydt,
                      n_store, B, D, H, W, cin, nout, act, alpha, _ptr(stats), stats.shape[0], _stream())
        elif ws is not None and wpack.shape[0] == 27:
            _lib.call("icsg3d_conv3d_k3_igemm_ws", _ptr(x), ldx, _ptr(wpack), _ptr(bias), _ptr(out), _ld(out), ydt, n_store,
                      B, D, H, W, cin, nout, act, alpha, _ptr(ws), ctypes.c_int64(ws.numel() * ws.element_size()), _stream())
        else:
            fn = "icsg3d_conv3d_k1_igemm" if wpack.shape[0] == 1 else "icsg3d_conv3d_k3_igemm"
            _lib.call(fn, _ptr(x), ldx, _ptr(wpack), _ptr(bias), _ptr(out), _ld(out), ydt, n_store, B, D, H, W, cin, nout,
                      act, alpha, _stream())
    return out


def pack_conv_w_dgrad_slice(w, ci0, cin, out=None):
    """dgrad operand bf16 [27][cin][cout] of the input-channel slice [ci0, ci0+cin) of w fp32 (3,3,3,cin_total,cout)."""
    _chk(w, torch.float32, "w")
    cin_total, cout = w.shape[-2], w.shape[-1]
    if out is None:
        out = torch.empty((27, cin, cout), dtype=torch.bfloat16, device=w.device)
    _lib.call("icsg3d_pack_conv_w_dgrad_slice", _ptr(w), _ptr(out), cin_total, ci0, cin, cout, _stream())
    return out


def pack_conv_w_upfold(w, c_skip0, c_skip, c_up0, c_up, out=None):
    """Keras kernel fp32 (3,3,3,cin,cout) -> the K-unit list of conv3d_k3_upfold: 27 taps of the skip channels
    [c_skip0, c_skip0+c_skip) and, per output phase, the 8 folded taps of the upsampled channels [c_up0, c_up0+c_up)."""
    _chk(w, torch.float32, "w")
    cin, cout = w.shape[-2], w.shape[-1]
    n = int(_lib.lib().icsg3d_conv3d_upfold_wpack_elems(c_skip, c_up, cout))
    if out is None:
        out = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    if out.numel() != n or not out.is_contiguous():
        raise ValueError("pack_conv_w_upfold: out has the wrong size")
    _lib.call("icsg3d_pack_conv_w_upfold", _ptr(w), cin, cout, c_skip0, c_skip, c_up0, c_up, _ptr(out), _stream())
    return out


def conv3d_k3_upfold(x_skip, x_low, wfold, bias, nout, *, act=ACT_NONE, alpha=LEAKY_ALPHA, out=None, post=None,
                     tag="conv.upfold", nominal=None):
    """Conv3D(3, same) over concatenate([x_skip, UpSampling3D(2)(x_low)]) without materialising the upsampled tensor
    (csrc/conv3d_upfold.cu).  x_skip bf16 [B,D,H,W,cs], x_low bf16 [B,D/2,H/2,W/2,cu] (channel slices allowed)."""
    _chk(x_skip, torch.bfloat16, "x_skip")
    _chk(x_low, torch.bfloat16, "x_low")
    _chk(wfold, torch.bfloat16, "wfold")
    B, D, H, W, cs = x_skip.shape
    cu = x_low.shape[-1]
    if tuple(x_low.shape[:4]) != (B, D // 2, H // 2, W // 2):
        raise ValueError("conv3d_k3_upfold: x_low must have half the extents of x_skip")
    if out is None:
        out = torch.empty((B, D, H, W, nout), dtype=torch.bfloat16, device=x_skip.device)
    _chk(out, torch.bfloat16, "out")
    ps, pt = post if post is not None else (None, None)
    vox = 2.0 * B * D * H * W * nout
    with _timed(("upfold", tag), vox * 27 * (nominal[0] if nominal else cs + cu), vox * (27 * cs + 8 * cu)):
        _lib.call("icsg3d_conv3d_k3_upfold", _ptr(x_skip), _ld(x_skip), cs, _ptr(x_low), _ld(x_low), cu, _ptr(wfold), _ptr(bias),
                  _ptr(ps), _ptr(pt), _ptr(out), _ld(out), out.shape[-1], B, D, H, W, nout, act, alpha, _stream())
    return out


def pack_conv_w_upfold_dgrad(w, c_up0, c_up, out=None):
    """Keras kernel fp32 (3,3,3,cin,cout) -> transposed folded weights for conv3d_k3_upfold_dgrad_low."""
    _chk(w, torch.float32, "w")
    cin, cout = w.shape[-2], w.shape[-1]
    n = int(_lib.lib().icsg3d_conv3d_upfold_dgrad_wpack_elems(cout, c_up))
    if out is None:
        out = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    if out.numel() != n or not out.is_contiguous():
        raise ValueError("pack_conv_w_upfold_dgrad: out has the wrong size")
    _lib.call("icsg3d_pack_conv_w_upfold_dgrad", _ptr(w), cin, cout, c_up0, c_up, _ptr(out), _stream())
    return out


def conv3d_k3_upfold_dgrad_low(dy, wpack, c_up, out=None, tag="conv.upfold.dgrad_low", nominal=None):
    """Gradient of conv3d_k3_upfold w.r.t. its low-resolution input: dy bf16 [B,D,H,W,cout] -> bf16 [B,D/2,H/2,W/2,c_up]."""
    _chk(dy, torch.bfloat16, "dy")
    _chk(wpack, torch.bfloat16, "wpack")
    B, D, H, W, cout = dy.shape
    if out is None:
        out = torch.empty((B, D // 2, H // 2, W // 2, c_up), dtype=torch.bfloat16, device=dy.device)
    _chk(out, torch.bfloat16, "out")
    vox = 2.0 * B * D * H * W * cout * c_up
    with _timed(("upfold", tag), vox * 27 if nominal is None else nominal, vox * 8):
        _lib.call("icsg3d_conv3d_k3_upfold_dgrad_low", _ptr(dy), _ld(dy), cout, _ptr(wpack), _ptr(out), _ld(out), c_up, B, D, H, W,
                  _stream())
    return out


def conv3d_k3_wgrad_workspace_bytes(B, D, cin, cout, k1=False):
    """Scratch bytes the filter-gradient kernel needs for this layer shape (split partials before the fixed-order sum)."""
    fn = _lib.lib().icsg3d_conv3d_k1_wgrad_workspace if k1 else _lib.lib().icsg3d_conv3d_k3_wgrad_workspace
    need = int(fn(B, D, D, D, cin, cout))
    if need < 0:
        raise _lib.Icsg3dError("conv3d wgrad workspace: invalid shape")
    return need


def _wgrad_scratch(need, ws, device):
    """Engines pass their own pre-sized workspace (one per stream; its address is baked into captured CUDA graphs, so it
    is never reallocated).  Direct callers without one get a fresh allocation per call — never a shared module-level
    buffer that a later, larger request could free under a captured graph."""
    if ws is None:
        return torch.empty(max(int(need), 16), dtype=torch.uint8, device=device)
    if ws.numel() * ws.element_size() < need:
        raise _lib.Icsg3dError(f"conv3d wgrad: workspace of {ws.numel() * ws.element_size()} B is smaller than the {need} B "
                               "this layer needs (size it with conv3d_k3_wgrad_workspace_bytes at construction)")
    return ws


def conv3d_k3_wgrad(x, dy, *, cin=None, cout=None, out=None, ref=False, nominal=None, tag="wgrad", ws=None):
    """dW[27][cin][cout] (fp32) = sum_v x[v+tap, ci] * dy[v, co]; x,dy bf16 NDHWC (channel slices allowed)."""
    _chk(x, torch.bfloat16, "x")
    _chk(dy, torch.bfloat16, "dy")
    B, D, H, W, _ = x.shape
    ldx, ldy = _ld(x), _ld(dy)
    cin = cin or x.shape[-1]
    cout = cout or dy.shape[-1]
    if out is None:
        out = torch.empty((27, cin, cout), dtype=torch.float32, device=x.device)
    if ref:
        _lib.call("icsg3d_ref_conv3d_k3_wgrad", _ptr(x), ldx, _ptr(dy), ldy, _ptr(out), B, D, H, W, cin, cout, _stream())
        return out
    need = _lib.lib().icsg3d_conv3d_k3_wgrad_workspace(B, D, H, W, cin, cout)
    if need < 0:
        raise _lib.Icsg3dError("conv3d_k3_wgrad_workspace: invalid shape")
    ws = _wgrad_scratch(need, ws, x.device)
    nc, no = nominal if nominal else (cin, cout)
    with _timed(("wgrad", tag), 2.0 * B * D * H * W * 27 * nc * no, 2.0 * B * D * H * W * 27 * cin * cout):
        _lib.call("icsg3d_conv3d_k3_wgrad", _ptr(x), ldx, _ptr(dy), ldy, _ptr(out), B, D, H, W, cin, cout, _ptr(ws),
                  ctypes.c_int64(ws.numel() * ws.element_size()), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# BatchNorm (+activation, +pool/upsample)
# ------------------------------------------------------------------------------------------------
BN_EPS = 1e-3
BN_MOMENTUM = 0.99


def _dt(t):
    return DT_BF16 if t.dtype == torch.bfloat16 else DT_F32


def bn_nparts(rows, C, dtype):
    n = _lib.lib().icsg3d_bn_nparts(ctypes.c_int64(rows), C, DT_BF16 if dtype == torch.bfloat16 else DT_F32)
    if n <= 0:
        raise _lib.Icsg3dError(f"bn_nparts: unsupported C={C} for {dtype}")
    return n


def bn_stats(x, C, partials):
    with _timed(("bn", "stats"), _nbytes(x)):
        return _bn_stats(x, C, partials)


def _bn_stats(x, C, partials):
    """x: [..., ld]; partials: float64 [nparts, 2, C]."""
    rows = x.numel() // x.shape[-1]
    _lib.call("icsg3d_bn_stats", _ptr(x), _ld(x), _dt(x), ctypes.c_int64(rows), C, _ptr(partials),
              partials.shape[0], _stream())


def bn_reduce_partials(partials, sums):
    _lib.call("icsg3d_bn_reduce_partials", _ptr(partials), partials.shape[0], partials.shape[2], _ptr(sums), _stream())


def bn_finalize(sums, count, gamma, beta, mean, rstd, scale, shift, moving_mean=None, moving_var=None, eps=BN_EPS,
                momentum=BN_MOMENTUM):
    C = mean.numel()
    _lib.call("icsg3d_bn_finalize", _ptr(sums), ctypes.c_double(count), _ptr(gamma), _ptr(beta), eps, _ptr(mean),
              _ptr(rstd), _ptr(scale), _ptr(shift), _ptr(moving_mean), _ptr(moving_var), momentum, C, _stream())


def bn_reduce_finalize(partials, count, gamma, beta, sums, mean, rstd, scale, shift, moving_mean=None, moving_var=None,
                       eps=BN_EPS, momentum=BN_MOMENTUM):
    C = mean.numel()
    _lib.call("icsg3d_bn_reduce_finalize", _ptr(partials), partials.shape[0], ctypes.c_double(count), _ptr(gamma), _ptr(beta),
              eps, _ptr(sums), _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift), _ptr(moving_mean), _ptr(moving_var),
              momentum, C, _stream())


def bn_reduce_grads(partials, sums, dgamma=None, dbeta=None):
    _lib.call("icsg3d_bn_reduce_grads", _ptr(partials), partials.shape[0], partials.shape[2], _ptr(sums), _ptr(dgamma),
              _ptr(dbeta), _stream())


def bn_allreduce_buffer_bytes(world, nslots, cmax):
    return int(_lib.lib().icsg3d_bn_allreduce_buffer_bytes(world, nslots, cmax))


def bn_reduce_allreduce_finalize(partials, count_global, gamma, beta, sums, mean, rstd, scale, shift, peers, world, rank, slot,
                                 nslots, cmax, epoch, moving_mean=None, moving_var=None, eps=BN_EPS, momentum=BN_MOMENTUM):
    """bn_reduce_finalize with the 2C sums all-reduced over NVLink peer memory inside the kernel (data parallel)."""
    _lib.call("icsg3d_bn_reduce_allreduce_finalize", _ptr(partials), partials.shape[0], ctypes.c_double(count_global),
              _ptr(gamma), _ptr(beta), eps, _ptr(sums), _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift), _ptr(moving_mean),
              _ptr(moving_var), momentum, partials.shape[2], _ptr(peers), world, rank, slot, nslots, cmax, _ptr(epoch),
              _stream())


def bn_reduce_allreduce_grads(partials, sums_global, peers, world, rank, slot, nslots, cmax, epoch, dgamma=None, dbeta=None):
    """bn_reduce_grads with the backward sums all-reduced over peer memory; dgamma/dbeta stay LOCAL sums."""
    _lib.call("icsg3d_bn_reduce_allreduce_grads", _ptr(partials), partials.shape[0], partials.shape[2], _ptr(sums_global),
              _ptr(dgamma), _ptr(dbeta), _ptr(peers), world, rank, slot, nslots, cmax, _ptr(epoch), _stream())


def bn_inference_coeffs(gamma, beta, moving_mean, moving_var, scale, shift, eps=BN_EPS):
    _lib.call("icsg3d_bn_inference_coeffs", _ptr(gamma), _ptr(beta), _ptr(moving_mean), _ptr(moving_var), eps,
              _ptr(scale), _ptr(shift), scale.numel(), _stream())


def bn_apply_fwd(x, C, scale, shift, act, post, y=None, y32=None, pool_idx=None, alpha=LEAKY_ALPHA):
    with _timed(("bn", "apply_fwd"), _nbytes(x, y, y32, pool_idx)):
        return _bn_apply_fwd(x, C, scale, shift, act, post, y, y32, pool_idx, alpha)


def _bn_apply_fwd(x, C, scale, shift, act, post, y=None, y32=None, pool_idx=None, alpha=LEAKY_ALPHA):
    B, D, H, W, _ = x.shape
    _lib.call("icsg3d_bn_apply_fwd", _ptr(x), _ld(x), _dt(x), _ptr(scale), _ptr(shift), act, alpha, post, B, D, H, W, C,
              _ptr(y), _ld(y) if y is not None else 0, _ptr(y32), y32.shape[-1] if y32 is not None else 0,
              _ptr(pool_idx), _stream())


# ---- fp32-class split operands (csrc/split3.cu) ----
SPLIT_FMT = 1  # fp32-class split operands: 0 = bf16 pairs, 1 = IEEE fp16 pairs (default: 22 significant bits)
SPLIT_WSCALE = 1024.0  # fp16 pairs: weights are packed x 2^10 (their lo parts stay normal numbers), the conv output x 2^-10


def _fmt(fmt):
    return SPLIT_FMT if fmt is None else fmt


def f32_to_split3(src, c, dst, ctot, coff=0, fmt=None):
    """src fp32 [..., ld] (first c channels) -> dst bf16 [..., 3*ctot] parts [hi | lo | hi] at channel offset coff."""
    rows = src.numel() // src.shape[-1]
    _lib.call("icsg3d_f32_to_split3", _ptr(src), src.shape[-1], c, ctypes.c_int64(rows), _ptr(dst), ctot, coff, _fmt(fmt),
              _stream())


def pack_vae_input_split3(m, cond, xe, xp, fmt=None):
    B = m.shape[0]
    vox = m.numel() // (B * 4)
    _lib.call("icsg3d_pack_vae_input_split3", _ptr(m), _ptr(cond), cond.shape[1] if cond is not None else 0, B,
              ctypes.c_int64(vox), _ptr(xe), _ptr(xp), _fmt(fmt), _stream())


def pack_vae_input_mixed(m, cond, xe3, xp16, fmt=0):
    """Split encoder operand + plain bf16 perceptual operand from one read of the fp32 batch."""
    B = m.shape[0]
    vox = m.numel() // (B * 4)
    _lib.call("icsg3d_pack_vae_input_mixed", _ptr(m), _ptr(cond), cond.shape[1], B, ctypes.c_int64(vox), _ptr(xe3),
              _ptr(xp16), fmt, _stream())


def pack_vae_input_lean(m, cond, xe32, xp16):
    """Lean 32-channel split encoder operand (see icsg3d_pack_vae_input_lean) + plain bf16 perceptual operand."""
    B = m.shape[0]
    vox = m.numel() // (B * 4)
    _lib.call("icsg3d_pack_vae_input_lean", _ptr(m), _ptr(cond), cond.shape[1], B, ctypes.c_int64(vox), _ptr(xe32), _ptr(xp16),
              _stream())


def pack_conv_w_dgrad_x3(w, cin_pad=None, cout_pad=None, out=None, fmt=None):
    """w fp32 (k,k,k,Cin,Cout) -> bf16 [taps][cin_pad][3*cout_pad] = [w_hi | w_hi | w_lo] along Cout, taps mirrored."""
    _chk(w, torch.float32, "w")
    ntaps = w.shape[0] * w.shape[1] * w.shape[2]
    cin, cout = w.shape[3], w.shape[4]
    cin_pad, cout_pad = cin_pad or pad16(cin), cout_pad or pad16(cout)
    if out is None:
        out = torch.empty((ntaps, cin_pad, 3 * cout_pad), dtype=torch.bfloat16, device=w.device)
    f = _fmt(fmt)
    _lib.call("icsg3d_pack_conv_w_dgrad_x3", _ptr(w), _ptr(out), ntaps, cin, cout, cin_pad, cout_pad, f,
              SPLIT_WSCALE if f == 1 else 1.0, _stream())
    return out


def conv3d_wgrad_x3(x3, dy3, cin_pad, cout_pad, out, k1=False, tag="wgrad.x3"):
    """Filter gradient from split tensors x3 [..., 3*cin_pad], dy3 [..., 3*cout_pad] (bf16 pairs): the ordinary wgrad
    kernel on the [hi | lo] channel slices, then the three-block sum -> out fp32 [taps][cin_pad][cout_pad]."""
    taps = 1 if k1 else 27
    P = torch.empty((taps, 2 * cin_pad, 2 * cout_pad), dtype=torch.float32, device=x3.device)
    xs, ds = x3[..., : 2 * cin_pad], dy3[..., : 2 * cout_pad]
    if k1:
        conv3d_k1_wgrad(xs, ds, cin=2 * cin_pad, cout=2 * cout_pad, out=P)
    else:
        conv3d_k3_wgrad(xs, ds, cin=2 * cin_pad, cout=2 * cout_pad, out=P, tag=tag)
    _lib.call("icsg3d_wgrad_combine_x3", _ptr(P), _ptr(out), taps, cin_pad, cout_pad, _stream())
    return out


def pack_conv_w_fprop_x3(w, cin_pad=None, cout_pad=None, cin_lead=0, fold=1, fold_c=0, out=None, fmt=None):
    """w fp32 (kd,kh,kw,Cin,Cout) (or (1,1,1,Cin,Cout)) -> bf16 [taps][cout_pad][3*cin_pad] = [w_hi | w_hi | w_lo]."""
    _chk(w, torch.float32, "w")
    ntaps = w.shape[0] * w.shape[1] * w.shape[2]
    cin, cout = w.shape[3], w.shape[4]
    if cin_pad is None:
        cin_pad = pad16(cin if fold <= 1 else cin_lead + fold_c)
    if cout_pad is None:
        cout_pad = pad16(cout)
    if out is None:
        out = torch.empty((ntaps, cout_pad, 3 * cin_pad), dtype=torch.bfloat16, device=w.device)
    f = _fmt(fmt)
    _lib.call("icsg3d_pack_conv_w_fprop_x3", _ptr(w), _ptr(out), ntaps, cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c,
              f, SPLIT_WSCALE if f == 1 else 1.0, _stream())
    return out


def bn_apply_fwd_split3(x, C, scale, shift, act, post, y, ctot, coff=0, pool_idx=None, alpha=LEAKY_ALPHA, fmt=None):
    B, D, H, W, _ = x.shape
    _lib.call("icsg3d_bn_apply_fwd_split3", _ptr(x), _ld(x), _dt(x), _ptr(scale), _ptr(shift), act, alpha, post, B, D, H, W, C,
              _ptr(y), _ld(y), _ptr(pool_idx), ctot, coff, _fmt(fmt), _stream())


def bn_bwd_nparts(x, C, post):
    B, D, H, W, _ = x.shape
    n = _lib.lib().icsg3d_bn_bwd_nparts(B, D, H, W, C, _dt(x), post)
    if n <= 0:
        raise _lib.Icsg3dError("bn_bwd_nparts: unsupported shape")
    return n


def bn_bwd_reduce(dy, x, C, mean, rstd, scale, shift, act, post, pool_idx, partials, alpha=LEAKY_ALPHA, dy2=None):
    with _timed(("bn", "bwd_reduce"), _nbytes(dy, x, pool_idx, dy2)):
        return _bn_bwd_reduce(dy, x, C, mean, rstd, scale, shift, act, post, pool_idx, partials, alpha, dy2)


def _bn_bwd_reduce(dy, x, C, mean, rstd, scale, shift, act, post, pool_idx, partials, alpha=LEAKY_ALPHA, dy2=None):
    B, D, H, W, _ = x.shape
    ldx = _ld(x)
    _lib.call("icsg3d_bn_bwd_reduce", _ptr(dy), _ld(dy), _ptr(dy2), _ld(dy2) if dy2 is not None else 0, _ptr(x), ldx, _dt(x), _ptr(mean), _ptr(rstd), _ptr(scale),
              _ptr(shift), act, alpha, post, _ptr(pool_idx), B, D, H, W, C, _ptr(partials), partials.shape[0], _stream())


def bn_bwd_apply_nblocks(x, C, post):
    """Partials written by bn_bwd_apply(..., tap_sq=...) for this layer shape."""
    B, D, H, W, _ = x.shape
    n = _lib.lib().icsg3d_bn_bwd_apply_nblocks(B, D, H, W, C, _dt(x), post)
    if n <= 0:
        raise _lib.Icsg3dError("bn_bwd_apply_nblocks: unsupported shape")
    return n


def bn_bwd_apply(dy, x, C, mean, rstd, scale, shift, act, post, pool_idx, sums, count, dx, pre_relu=False,
                 tap_other=None, tap_coef=0.0, alpha=LEAKY_ALPHA, dy2=None, tap_sq=None):
    """tap_sq (fp64 [bn_bwd_apply_nblocks]): also return the per-block sums (x - tap_other)^2 of a tapped layer."""
    with _timed(("bn", "bwd_apply"), _nbytes(dy, x, pool_idx, dy2, dx, tap_other)):
        if tap_sq is not None:
            assert dy2 is None and tap_other is not None and tap_sq.dtype == torch.float64
            B, D, H, W, _ = x.shape
            return _lib.call("icsg3d_bn_bwd_apply_tapsq", _ptr(dy), _ld(dy), _ptr(x), _ld(x), _dt(x), _ptr(mean), _ptr(rstd),
                             _ptr(scale), _ptr(shift), act, alpha, post, _ptr(pool_idx), B, D, H, W, C, _ptr(sums),
                             ctypes.c_double(count), 1 if pre_relu else 0, _ptr(tap_other), tap_other.shape[-1], tap_coef,
                             _ptr(dx), _ld(dx), _ptr(tap_sq), tap_sq.numel(), _stream())
        return _bn_bwd_apply(dy, x, C, mean, rstd, scale, shift, act, post, pool_idx, sums, count, dx, pre_relu, tap_other,
                             tap_coef, alpha, dy2)


def _bn_bwd_apply(dy, x, C, mean, rstd, scale, shift, act, post, pool_idx, sums, count, dx, pre_relu=False,
                  tap_other=None, tap_coef=0.0, alpha=LEAKY_ALPHA, dy2=None):
    B, D, H, W, _ = x.shape
    ldx = _ld(x)
    _lib.call("icsg3d_bn_bwd_apply", _ptr(dy), _ld(dy), _ptr(dy2), _ld(dy2) if dy2 is not None else 0, _ptr(x), ldx, _dt(x), _ptr(mean), _ptr(rstd), _ptr(scale),
              _ptr(shift), act, alpha, post, _ptr(pool_idx), B, D, H, W, C, _ptr(sums), ctypes.c_double(count),
              1 if pre_relu else 0, _ptr(tap_other), tap_other.shape[-1] if tap_other is not None else 0,
              tap_coef, _ptr(dx), _ld(dx), _stream())


def bn_bwd_apply_f32(dy, x, C, mean, rstd, scale, shift, act, post, pool_idx, sums, count, dx, pre_relu=False,
                     tap_other=None, tap_coef=0.0, alpha=LEAKY_ALPHA, dy2=None):
    """bn_bwd_apply with every tensor in fp32 (fp32-class backward): x, dy, dy2, tap_other, dx."""
    for t, n in ((dy, "dy"), (x, "x"), (dx, "dx"), (dy2, "dy2"), (tap_other, "tap_other")):
        if t is not None:
            _chk(t, torch.float32, n)
    B, D, H, W, _ = x.shape
    _lib.call("icsg3d_bn_bwd_apply_f32", _ptr(dy), _ld(dy), _ptr(dy2), _ld(dy2) if dy2 is not None else 0, _ptr(x), _ld(x),
              _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift), act, alpha, post, _ptr(pool_idx), B, D, H, W, C, _ptr(sums),
              ctypes.c_double(count), 1 if pre_relu else 0, _ptr(tap_other), _ld(tap_other) if tap_other is not None else 0,
              tap_coef, _ptr(dx), _ld(dx), _stream())


def xhat_grad_f32(x, xhat, mse_coef, dpm, dy):
    rows = x.numel() // 4
    _lib.call("icsg3d_xhat_grad_f32", _ptr(x), _ptr(xhat), mse_coef, _ptr(dpm), dpm.shape[-1] if dpm is not None else 0,
              ctypes.c_int64(rows), _ptr(dy), _stream())


def tap_grad_relu_f32(a, other, coef, dc):
    _lib.call("icsg3d_tap_grad_relu_f32", _ptr(a), _ptr(other), coef, ctypes.c_int64(a.numel()), _ptr(dc), _stream())


def act_bwd_f32(dy, y, act, dx, alpha=LEAKY_ALPHA):
    _lib.call("icsg3d_act_bwd_f32", _ptr(dy), _ptr(y), act, alpha, ctypes.c_int64(y.numel()), _ptr(dx), _stream())


def bn_bwd_fused_nparts(C, dtype):
    n = _lib.lib().icsg3d_bn_bwd_fused_nparts(C, DT_BF16 if dtype == torch.bfloat16 else DT_F32)
    if n <= 0:
        raise _lib.Icsg3dError("bn_bwd_fused_nparts: unsupported shape")
    return n


def bn_bwd_fused(dy, x, C, mean, rstd, scale, shift, act, post, pool_idx, partials, sums, count_global, dx, dgamma=None,
                 dbeta=None, pre_relu=False, tap_other=None, tap_coef=0.0, alpha=LEAKY_ALPHA, dy2=None, peers=None, world=1,
                 rank=0, slot=0, nslots=1, cmax=0, epoch=None):
    """Whole BatchNorm backward of one layer in one cooperative launch (reduce -> [peer all-reduce] -> apply)."""
    B, D, H, W, _ = x.shape
    _lib.call("icsg3d_bn_bwd_fused", _ptr(dy), _ld(dy), _ptr(dy2), _ld(dy2) if dy2 is not None else 0, _ptr(x), _ld(x), _dt(x),
              _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift), act, alpha, post, _ptr(pool_idx), B, D, H, W, C, _ptr(partials),
              _ptr(sums), ctypes.c_double(count_global), _ptr(dgamma), _ptr(dbeta), 1 if pre_relu else 0, _ptr(tap_other),
              tap_other.shape[-1] if tap_other is not None else 0, tap_coef, _ptr(dx), _ld(dx), _ptr(peers), world, rank, slot,
              nslots, cmax, _ptr(epoch), _stream())


def bn_param_grads(sums, dgamma, dbeta):
    C = sums.numel() // 2
    _lib.call("icsg3d_bn_param_grads", _ptr(sums), _ptr(dgamma), _ptr(dbeta), C, _stream())


# ------------------------------------------------------------------------------------------------
# VAE glue
# ------------------------------------------------------------------------------------------------
def pack_vae_input(m, cond, xe, xp):
    B = m.shape[0]
    vox = m.numel() // (B * 4)
    _lib.call("icsg3d_pack_vae_input", _ptr(m), _ptr(cond), cond.shape[1] if cond is not None else 0, B,
              ctypes.c_int64(vox), _ptr(xe), _ptr(xp), _stream())


def f32_to_bf16_rows(src, c, dst):
    rows = src.numel() // c
    _lib.call("icsg3d_f32_to_bf16_rows", _ptr(src), c, ctypes.c_int64(rows), _ptr(dst), dst.shape[-1], _stream())


def bf16_rows_to_f32(src, c, dst):
    rows = src.numel() // src.shape[-1]
    _lib.call("icsg3d_bf16_rows_to_f32", _ptr(src), src.shape[-1], c, ctypes.c_int64(rows), _ptr(dst), _stream())


def dense_fwd(x1, w, bias, y, act=ACT_NONE, x2=None):
    B = x1.shape[0]
    k1 = x1.numel() // B
    k2 = x2.numel() // B if x2 is not None else 0
    _lib.call("icsg3d_dense_fwd", _ptr(x1), k1, _ptr(x2), k2, _ptr(w), _ptr(bias), act, B, w.shape[1], _ptr(y), _stream())


def dense_bwd(x1, w, y, dy, dw, db, act=ACT_NONE, x2=None, dx1=None, accumulate_dx=False):
    B = x1.shape[0]
    k1 = x1.numel() // B
    k2 = x2.numel() // B if x2 is not None else 0
    _lib.call("icsg3d_dense_bwd", _ptr(x1), k1, _ptr(x2), k2, _ptr(w), _ptr(y), act, _ptr(dy), B, w.shape[1], _ptr(dx1),
              1 if accumulate_dx else 0, _ptr(dw), _ptr(db), _stream())


def reparam_fwd(mu, lv, eps, z, kl):
    _lib.call("icsg3d_reparam_fwd", _ptr(mu), _ptr(lv), _ptr(eps), mu.shape[0], mu.shape[1], _ptr(z), _ptr(kl), _stream())


def reparam_bwd(dz, mu, lv, eps, kl_coef, dmu, dlv):
    _lib.call("icsg3d_reparam_bwd", _ptr(dz), _ptr(mu), _ptr(lv), _ptr(eps), kl_coef, mu.shape[0], mu.shape[1], _ptr(dmu),
              _ptr(dlv), _stream())


def leaky_bwd_rows(dy, y, c, dst, alpha=LEAKY_ALPHA):
    rows = y.numel() // c
    _lib.call("icsg3d_leaky_bwd_rows", _ptr(dy), _ptr(y), alpha, c, ctypes.c_int64(rows), _ptr(dst), dst.shape[-1],
              _stream())


def sqdiff_nparts(n):
    return _lib.lib().icsg3d_sqdiff_nparts(ctypes.c_int64(n))


def sqdiff_partials(a, b, partials, nparts):
    """partials: float64 view with at least nparts entries."""
    _lib.call("icsg3d_sqdiff_partials", _ptr(a), _ptr(b), _dt(a), ctypes.c_int64(a.numel()), _ptr(partials), nparts,
              _stream())


def vae_loss_assemble(partials, nparts, scales, kl, kl_scale, alpha, beta, out, raw=None):
    """partials: float64 [nterms, stride]; nparts: int32 [nterms]; scales: float64 [nterms]."""
    _lib.call("icsg3d_vae_loss_assemble", _ptr(partials), _ptr(nparts), partials.shape[1], partials.shape[0],
              _ptr(scales), _ptr(kl), kl.numel(), ctypes.c_double(kl_scale), alpha, beta, _ptr(out), _ptr(raw), _stream())


def xhat_grad(x, xhat, mse_coef, dpm, dy):
    rows = x.numel() // 4
    _lib.call("icsg3d_xhat_grad", _ptr(x), _ptr(xhat), mse_coef, _ptr(dpm), dpm.shape[-1] if dpm is not None else 0,
              ctypes.c_int64(rows), _ptr(dy), _stream())


def tap_grad_relu(a, other, coef, dc):
    _lib.call("icsg3d_tap_grad_relu", _ptr(a), _ptr(other), coef, ctypes.c_int64(a.numel()), _ptr(dc), _stream())


def bias_grad(dy, C, db):
    rows = dy.numel() // dy.shape[-1]
    _lib.call("icsg3d_bias_grad", _ptr(dy), dy.shape[-1], ctypes.c_int64(rows), C, _ptr(db), _stream())


def adam_keras_step(p, g, m, v, state, lr, beta1=0.9, beta2=0.999, eps=1e-7, grad_scale=1.0):
    _lib.call("icsg3d_adam_keras_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(state), ctypes.c_double(lr),
              ctypes.c_double(beta1), ctypes.c_double(beta2), ctypes.c_double(eps), grad_scale,
              ctypes.c_int64(p.numel()), _stream())


def adam_allreduce_buffer_bytes(world, n):
    return int(_lib.lib().icsg3d_adam_allreduce_buffer_bytes(world, ctypes.c_int64(n)))


def adam_keras_allreduce_step(p, g, m, v, state, lr, peers, world, rank, epoch, beta1=0.9, beta2=0.999, eps=1e-7,
                              grad_scale=1.0):
    """Gradient all-reduce over NVLink peer memory fused with Keras-Adam (data parallel); g becomes the global sum."""
    _lib.call("icsg3d_adam_keras_allreduce_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(state), ctypes.c_double(lr),
              ctypes.c_double(beta1), ctypes.c_double(beta2), ctypes.c_double(eps), grad_scale, ctypes.c_int64(p.numel()),
              _ptr(peers), world, rank, _ptr(epoch), _stream())


# ------------------------------------------------------------------------------------------------
# U-Net heads
# ------------------------------------------------------------------------------------------------
def conv3d_k1_wgrad(x, dy, *, cin, cout, out, ws=None):
    """dW[1][cin][cout] of a 1x1x1 conv (tcgen05, same kernel as the 3x3x3 filter gradient)."""
    B, D, H, W, _ = x.shape
    need = _lib.lib().icsg3d_conv3d_k1_wgrad_workspace(B, D, H, W, cin, cout)
    ws = _wgrad_scratch(need, ws, x.device)
    with _timed(("wgrad", "heads.wgrad"), 2.0 * B * D * H * W * cin * cout):
        _lib.call("icsg3d_conv3d_k1_wgrad", _ptr(x), _ld(x), _ptr(dy), _ld(dy), _ptr(out), B, D, H, W, cin, cout, _ptr(ws),
                  ctypes.c_int64(ws.numel() * ws.element_size()), _stream())
    return out


def pack_heads_w(w_soft, w_sig, b_soft, b_sig, wf, wd, bias):
    cin, c1 = w_soft.shape[-2], w_soft.shape[-1]
    _lib.call("icsg3d_pack_heads_w", _ptr(w_soft), _ptr(w_sig), _ptr(b_soft), _ptr(b_sig), cin, c1, wf.shape[1], _ptr(wf),
              _ptr(wd), _ptr(bias), _stream())


def unpack_heads_grad(dwcat, colsum, c1, dw_soft, dw_sig, db_soft, db_sig):
    cin, nout = dwcat.shape[-2], dwcat.shape[-1]
    _lib.call("icsg3d_unpack_heads_grad", _ptr(dwcat), _ptr(colsum), cin, c1, nout, _ptr(dw_soft), _ptr(dw_sig),
              _ptr(db_soft), _ptr(db_sig), _stream())


def heads_loss_nparts(M):
    return _lib.lib().icsg3d_heads_loss_nparts(ctypes.c_int64(M))


def heads_loss(logits, c1, species, class_w, inv_count, partials, argmax_out=None, sig_prob=None, dlogits=None, probs=None):
    M = logits.numel() // logits.shape[-1]
    if dlogits is not None and dlogits.dtype == torch.float32:  # fp32-class backward
        return _lib.call("icsg3d_heads_loss_f32grad", _ptr(logits), logits.shape[-1], c1, _ptr(species), _ptr(class_w),
                         ctypes.c_int64(M), inv_count, _ptr(argmax_out), _ptr(sig_prob), _ptr(probs), _ptr(dlogits),
                         dlogits.shape[-1], _ptr(partials), partials.shape[0], _stream())
    _lib.call("icsg3d_heads_loss", _ptr(logits), logits.shape[-1], c1, _ptr(species), _ptr(class_w), ctypes.c_int64(M),
              inv_count, _ptr(argmax_out), _ptr(sig_prob), _ptr(probs), _ptr(dlogits), dlogits.shape[-1] if dlogits is not None else 0,
              _ptr(partials), partials.shape[0], _stream())


def heads_loss_fused_nparts(M):
    return _lib.lib().icsg3d_heads_loss_fused_nparts(ctypes.c_int64(M))


def heads_loss_fused(x, wpack, bias, c1, species, class_w, inv_count, partials, argmax_out=None, sig_prob=None, dlogits=None):
    """Head GEMM + both losses + metric counts + bf16 d(loss)/d(logits) in one kernel (csrc/heads_fused.cu); `partials`
    fp64 [>= heads_loss_fused_nparts(M), 6]: only the first nparts rows are written — pass exactly that view to
    heads_loss_finalize."""
    _chk(x, torch.bfloat16, "x")
    _chk(wpack, torch.bfloat16, "wpack")
    if dlogits is not None:
        _chk(dlogits, torch.bfloat16, "dlogits")
    nout, cin = wpack.shape[-2], wpack.shape[-1]
    M = x.numel() // x.shape[-1]
    n = heads_loss_fused_nparts(M)
    if partials.shape[0] != n or not partials.is_contiguous():
        raise ValueError(f"heads_loss_fused: partials must be a contiguous [{n}, 6] fp64 tensor")
    _lib.call("icsg3d_heads_loss_fused", _ptr(x), x.stride(-2), _ptr(wpack), _ptr(bias), ctypes.c_int64(M), cin, nout, c1,
              _ptr(species), _ptr(class_w), inv_count, _ptr(partials), _ptr(argmax_out), _ptr(sig_prob), _ptr(dlogits),
              dlogits.shape[-1] if dlogits is not None else 0, _stream())


def heads_loss_finalize(partials, count, out, raw=None):
    _lib.call("icsg3d_heads_loss_finalize", _ptr(partials), partials.shape[0], ctypes.c_double(count), _ptr(out), _ptr(raw),
              _stream())


def heads_predict(logits, c1, threshold, argmax=None, mask=None, sig_prob=None):
    """generate.py:221-225 on the fp32 head logits: argmax species (uint8), sigmoid >= threshold mask (uint8), sigmoid."""
    _chk(logits, torch.float32, "logits")
    M = logits.numel() // logits.shape[-1]
    _lib.call("icsg3d_heads_predict", _ptr(logits), logits.shape[-1], c1, ctypes.c_int64(M), float(threshold), _ptr(argmax),
              _ptr(mask), _ptr(sig_prob), _stream())


def heads_predict_fused(x, wpack, bias, c1, threshold, argmax=None, mask=None, sig_prob=None, cin=None, f16=False,
                        out_scale=1.0):
    """generate.py:220-225 without the logits round trip: head GEMM on the features x [..., ld] (first `cin` channels)
    with the packed head weights [1][nout][cin], soft-max arg-max + sigmoid threshold in the GEMM's epilogue."""
    _chk(x, torch.bfloat16, "x")
    _chk(wpack, torch.bfloat16, "wpack")
    nout, k = wpack.shape[-2], wpack.shape[-1]
    cin = k if cin is None else cin
    if cin != k:
        raise ValueError("heads_predict_fused: wpack does not match cin")
    M = x.numel() // x.shape[-1]
    _lib.call("icsg3d_heads_predict_fused", _ptr(x), x.stride(-2), _ptr(wpack), _ptr(bias), ctypes.c_int64(M), cin, nout, c1,
              1 if f16 else 0, float(out_scale), float(threshold), _ptr(argmax), _ptr(mask), _ptr(sig_prob), _stream())


def metric_counts(y_true, y_pred):
    """unet.py:159-193: the five K.round(K.clip(.)) sums over (y_true, y_pred) fp32 [..., C] -> fp64 [5] device tensor."""
    _chk(y_true, torch.float32, "y_true")
    _chk(y_pred, torch.float32, "y_pred")
    if y_true.shape != y_pred.shape or not (y_true.is_contiguous() and y_pred.is_contiguous()):
        raise ValueError("metric_counts: y_true / y_pred must be contiguous tensors of the same shape")
    counts = torch.empty(5, dtype=torch.float64, device=y_true.device)
    _lib.call("icsg3d_metric_counts", _ptr(y_true), _ptr(y_pred), ctypes.c_int64(y_true.numel()), y_true.shape[-1],
              _ptr(counts), _stream())
    return counts
