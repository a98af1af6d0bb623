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


def _chk(t, dtype, name):
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")


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


def unpack_conv_dw(dw_pad, cin, cout, cin_lead=0, fold=1, fold_c=0, out=None):
    _chk(dw_pad, torch.float32, "dw_pad")
    cin_pad, cout_pad = dw_pad.shape[1], dw_pad.shape[2]
    if out is None:
        out = torch.empty((3, 3, 3, cin, cout), dtype=torch.float32, device=dw_pad.device)
    _lib.call("icsg3d_unpack_conv_dw", _ptr(dw_pad), _ptr(out), cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c,
              _stream())
    return out


def conv3d_k3(x, wpack, bias=None, *, cin=None, n_store=None, act=ACT_NONE, alpha=LEAKY_ALPHA, out=None,
              out_dtype=torch.bfloat16, ref=False):
    """x: bf16 [B,D,H,W,ldx]; wpack: bf16 [27][nout][cin]; returns y [B,D,H,W,n_store].

    `ref=True` runs the CUDA-core cross-check kernel (fp32 output) instead of the tcgen05 kernel.
    """
    _chk(x, torch.bfloat16, "x")
    _chk(wpack, torch.bfloat16, "wpack")
    B, D, H, W, ldx = x.shape
    nout, cin_w = wpack.shape[1], wpack.shape[2]
    cin = cin or cin_w
    if cin != cin_w:
        raise ValueError(f"conv3d_k3: cin {cin} does not match packed weights {cin_w}")
    n_store = n_store or nout
    if bias is not None:
        _chk(bias, torch.float32, "bias")
        if bias.numel() < nout:
            raise ValueError("bias must be padded to nout")
    if ref:
        y = torch.empty((B, D, H, W, nout), dtype=torch.float32, device=x.device) if out is None else out
        _lib.call("icsg3d_ref_conv3d_k3", _ptr(x), ldx, _ptr(wpack), _ptr(bias), _ptr(y), y.shape[-1], B, D, H, W, cin,
                  nout, act, alpha, _stream())
        return y
    if out is None:
        out = torch.empty((B, D, H, W, n_store), dtype=out_dtype, device=x.device)
    ydt = DT_BF16 if out.dtype == torch.bfloat16 else DT_F32
    _lib.call("icsg3d_conv3d_k3_igemm", _ptr(x), ldx, _ptr(wpack), _ptr(bias), _ptr(out), out.shape[-1], ydt, n_store,
              B, D, H, W, cin, nout, act, alpha, _stream())
    return out


_wgrad_ws = {}


def conv3d_k3_wgrad(x, dy, *, cin=None, cout=None, out=None, ref=False):
    """dW[27][cin][cout] (fp32) = sum_v x[v+tap, ci] * dy[v, co]; x,dy bf16 NDHWC."""
    _chk(x, torch.bfloat16, "x")
    _chk(dy, torch.bfloat16, "dy")
    B, D, H, W, ldx = x.shape
    ldy = dy.shape[-1]
    cin = cin or ldx
    cout = cout or ldy
    if out is None:
        out = torch.empty((27, cin, cout), dtype=torch.float32, device=x.device)
    if ref:
        _lib.call("icsg3d_ref_conv3d_k3_wgrad", _ptr(x), ldx, _ptr(dy), ldy, _ptr(out), B, D, H, W, cin, cout, _stream())
        return out
    need = _lib.lib().icsg3d_conv3d_k3_wgrad_workspace(B, D, H, W, cin, cout)
    if need < 0:
        raise _lib.Icsg3dError("conv3d_k3_wgrad_workspace: invalid shape")
    key = (x.device.index,)
    ws = _wgrad_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(int(need), dtype=torch.uint8, device=x.device)
        _wgrad_ws[key] = ws
    _lib.call("icsg3d_conv3d_k3_wgrad", _ptr(x), ldx, _ptr(dy), ldy, _ptr(out), B, D, H, W, cin, cout, _ptr(ws),
              ctypes.c_int64(ws.numel()), _stream())
    return out
