"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement, in plain PyTorch, of the Keras 2.3.1 / TF 2.1 layer semantics the reference graphs use
(SURVEY.md §8c "[K] semantics register" R3-R8).  Tensors are channels-last NDHWC like the reference;
weights keep the Keras layouts (Conv3D kernel (kd,kh,kw,Cin,Cout), Dense kernel (in,out)).

Parity status: UNPINNED for the networks.  Keras/TensorFlow are not installable here and the reference
ships no tests, golden vectors or weights, so these functions restate the published layer
definitions; they are pinned only structurally (parameter counts, shapes, loss constants — see
tests/test_oracle_networks.py) and by fp64 gradient checks.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-3        # keras.layers.BatchNormalization default epsilon (R3)
BN_MOMENTUM = 0.99   # default momentum (R3)
LEAKY_ALPHA = 0.3    # keras.layers.LeakyReLU default alpha (R4)


def conv3d_same(x, kernel, bias=None):
    """Keras Conv3D(padding="same", strides 1): cross-correlation, NDHWC, kernel (kd,kh,kw,Cin,Cout) (R5).
    Used at lattice_vae.py:173,178,213,219 and unet.py:276-336."""
    kd = kernel.shape[0]
    w = kernel.permute(4, 3, 0, 1, 2)
    y = F.conv3d(x.permute(0, 4, 1, 2, 3), w, bias, padding=kd // 2)
    return y.permute(0, 2, 3, 4, 1)


# Data-parallel restatement hook (tests only): when set to a differentiable all-reduce-sum, batch statistics are
# computed from (sum x, sum x^2, count) summed over the ranks — the sync-BN scheme of SURVEY §8e — so that N ranks on
# batch shards reproduce one rank on the full batch.
STAT_ALLREDUCE = None


def batchnorm(x, gamma, beta, moving_mean, moving_var, training, eps=BN_EPS):
    """keras BatchNormalization(axis=-1) on a 5-D tensor (R3): biased batch variance in training,
    moving statistics otherwise.  Returns (y, batch_mean, batch_var) (stats are None in inference)."""
    if training and STAT_ALLREDUCE is not None:
        dims = tuple(range(x.dim() - 1))
        n_local = x.numel() // x.shape[-1]
        packed = torch.cat([x.sum(dim=dims), (x * x).sum(dim=dims), torch.full((1,), float(n_local), dtype=x.dtype)])
        packed = STAT_ALLREDUCE(packed)
        C = x.shape[-1]
        n = packed[-1]
        mean = packed[:C] / n
        var = packed[C:2 * C] / n - mean * mean
        y = (x - mean) / torch.sqrt(var + eps) * gamma + beta
        return y, mean, var
    if training:
        dims = tuple(range(x.dim() - 1))
        mean = x.mean(dim=dims)
        var = ((x - mean) ** 2).mean(dim=dims)
        y = (x - mean) / torch.sqrt(var + eps) * gamma + beta
        return y, mean, var
    y = (x - moving_mean) / torch.sqrt(moving_var + eps) * gamma + beta
    return y, None, None


def bn_moving_update(moving_mean, moving_var, mean, var, n, momentum=BN_MOMENTUM, eps=BN_EPS):
    """Keras 2.3.1 TF-backend moving-average update: mov -= (mov - val) * (1 - momentum), with the variance
    scaled by n / (n - (1 + eps)) (R3)."""
    var_unbiased = var * (n / (n - (1.0 + eps)))
    new_mean = moving_mean - (moving_mean - mean) * (1.0 - momentum)
    new_var = moving_var - (moving_var - var_unbiased) * (1.0 - momentum)
    return new_mean, new_var


def leaky_relu(x, alpha=LEAKY_ALPHA):
    """LeakyReLU(): gradient at exactly 0 is alpha (R4)."""
    return torch.where(x > 0, x, alpha * x)


def relu(x):
    """ReLU(): gradient at exactly 0 is 0 (R4)."""
    return torch.where(x > 0, x, torch.zeros_like(x))


def maxpool2(x, return_index=False):
    """MaxPool3D(2) / MaxPool3D(strides=2): window 2, stride 2, valid (R6).
    Backward tie rule: the FIRST maximum in (d,h,w) scan order receives the whole gradient (R6 decision)."""
    B, D, H, W, C = x.shape
    xs = x.reshape(B, D // 2, 2, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 5, 7, 2, 4, 6)
    xs = xs.reshape(B, D // 2, H // 2, W // 2, C, 8)  # last axis = (dd,dh,dw) scan order
    mx = xs.max(dim=-1, keepdim=True).values
    eq = xs == mx
    first = eq & (eq.cumsum(dim=-1) == 1)
    idx = first.to(torch.uint8).argmax(dim=-1)
    y = (xs * first.to(xs.dtype)).sum(dim=-1)
    # `first` is a constant mask for autograd: gradient flows only to the selected element.
    if return_index:
        return y, idx
    return y


def upsample2(x):
    """UpSampling3D(2): nearest-neighbour repeat (R7)."""
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def upsample2_conv3d_same_folded(x_low, kernel, bias=None):
    """conv3d_same(upsample2(x_low), kernel) restated on the LOW-resolution tensor (the identity behind
    csrc/conv3d_upfold.cu, used where unet.py:309-332 and lattice_vae.py:211-217 put a Conv3D after UpSampling3D).
    Per axis and output parity r (o = 2i + r) the three taps k = -1, 0, +1 land on low index i + floor((r + k) / 2):
        r = 0:  {-1} -> i-1,  {0, +1} -> i          r = 1:  {-1, 0} -> i,  {+1} -> i+1
    so every one of the 8 output phases is a 2x2x2 cross-correlation with summed taps (offsets {-1, 0} for r = 0,
    {0, +1} for r = 1); zero padding carries over because low index -1 / D is exactly upsampled index -1 / 2D."""
    B, D, H, W, _ = x_low.shape
    cout = kernel.shape[-1]
    y = x_low.new_zeros(B, 2 * D, 2 * H, 2 * W, cout)
    sets = {0: ([0], [1, 2]), 1: ([0, 1], [2])}          # parity -> kernel indices folded onto low tap t = 0, 1
    xp = F.pad(x_low.permute(0, 4, 1, 2, 3), (1, 1, 1, 1, 1, 1))   # one low-resolution voxel of zero padding per side
    for rd in (0, 1):
        for rh in (0, 1):
            for rw in (0, 1):
                wf = x_low.new_zeros(2, 2, 2, kernel.shape[3], cout)
                for td in (0, 1):
                    for th in (0, 1):
                        for tw in (0, 1):
                            for kd in sets[rd][td]:
                                for kh in sets[rh][th]:
                                    for kw in sets[rw][tw]:
                                        wf[td, th, tw] += kernel[kd, kh, kw]
                # low offsets t - (1 - r): the padded tensor starts at -1, so the window of phase r starts at index r
                win = xp[:, :, rd:rd + D + 1, rh:rh + H + 1, rw:rw + W + 1]
                ph = F.conv3d(win, wf.permute(4, 3, 0, 1, 2))
                y[:, rd::2, rh::2, rw::2, :] = ph.permute(0, 2, 3, 4, 1)
    return y if bias is None else y + bias


def dense(x, kernel, bias):
    """Dense: kernel (in,out) (R8)."""
    return x @ kernel + bias


def glorot_uniform(shape, fan_in, fan_out, gen):
    """keras glorot_uniform: U(-l, l), l = sqrt(6 / (fan_in + fan_out)) (R5, R8)."""
    limit = (6.0 / (fan_in + fan_out)) ** 0.5
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit
