"""T1 kernel-local parity for the fused BatchNorm(+activation)(+MaxPool3D/UpSampling3D) passes, forward and
backward, against the oracle's Keras-semantics ops differentiated by torch autograd (fp64)."""
import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu

CASES = [  # (B, D, C, act, post, order)   order: "vae" = BN->act->post ; "unet" = ReLU(x) is the BN input
    (2, 8, 16, "leaky", "pool", "vae"), (2, 8, 64, "leaky", "up", "vae"), (2, 8, 32, "leaky", "none", "vae"),
    (2, 8, 64, "none", "pool", "unet"), (2, 4, 512, "none", "none", "unet"), (3, 4, 128, "none", "pool", "unet"),
]


def _oracle(x, gamma, beta, act, post, order, dy, tap_other=None, tap_coef=0.0):
    from oracle import keras_ops as K
    x = x.double().requires_grad_(True)
    a = K.relu(x) if order == "unet" else x
    y, mean, var = K.batchnorm(a, gamma.double(), beta.double(), None, None, True)
    if act == "leaky":
        y = K.leaky_relu(y)
    elif act == "relu":
        y = K.relu(y)
    if post == "pool":
        y = K.maxpool2(y)
    elif post == "up":
        y = K.upsample2(y)
    loss = (y * dy.double()).sum()
    if tap_other is not None:
        loss = loss + 0.5 * tap_coef * ((a - tap_other.double()) ** 2).sum()
    loss.backward()
    return y.detach(), x.grad, mean.detach(), var.detach()


@pytest.mark.parametrize("B,D,C,act,post,order", CASES)
def test_bn_fwd_bwd(B, D, C, act, post, order):
    from icsg3d_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(B, D, D, D, C, generator=g) * 1.5 + 0.7)
    if order == "unet":
        x = torch.relu(x)  # the kernel sees the post-ReLU activation, as written by the conv epilogue
    x = x.to(torch.bfloat16)
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g) * 0.1
    Do = D // 2 if post == "pool" else D * 2 if post == "up" else D
    dy = torch.randn(B, Do, Do, Do, C, generator=g).to(torch.bfloat16)
    tap_other = torch.relu(torch.randn(B, D, D, D, C, generator=g)).to(torch.bfloat16) if order == "unet" else None
    tap_coef = 0.37 if order == "unet" else 0.0
    xin = x.float()
    if order == "unet":
        # oracle differentiates through the ReLU; feed it a pre-activation that reproduces x (>0 kept, 0 -> -1)
        xin = torch.where(x.float() > 0, x.float(), torch.full_like(x.float(), -1.0))
    y_ref, dx_ref, mean_ref, var_ref = _oracle(xin, gamma, beta, act, post, order, dy.float(), tap_other, tap_coef)

    dev = "cuda"
    A = {"none": ops.ACT_NONE, "leaky": ops.ACT_LEAKY, "relu": ops.ACT_RELU}[act]
    P = {"none": ops.POST_NONE, "pool": ops.POST_POOL2, "up": ops.POST_UP2}[post]
    xd, dyd = x.to(dev), dy.to(dev)
    rows = B * D ** 3
    n = ops.bn_nparts(rows, C, torch.bfloat16)
    part = torch.zeros(n, 2, C, dtype=torch.float64, device=dev)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    mean, rstd, scale, shift = (torch.zeros(C, device=dev) for _ in range(4))
    mm, mv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ops.bn_stats(xd, C, part)
    ops.bn_reduce_partials(part, sums)
    ops.bn_finalize(sums, float(rows), gamma.to(dev), beta.to(dev), mean, rstd, scale, shift, mm, mv)
    y = torch.zeros(B, Do, Do, Do, C, dtype=torch.bfloat16, device=dev)
    idx = torch.zeros(B, Do, Do, Do, C, dtype=torch.uint8, device=dev) if post == "pool" else None
    ops.bn_apply_fwd(xd, C, scale, shift, A, P, y=y, pool_idx=idx)
    torch.cuda.synchronize()
    assert rel_l2(mean, mean_ref) < 1e-5
    assert rel_l2(1.0 / rstd.double() ** 2 - 1e-3, var_ref) < 1e-4
    assert rel_l2(y.float(), y_ref) < 1e-2
    # Keras moving-average update
    from oracle import keras_ops as K
    emm, emv = K.bn_moving_update(torch.zeros(C).double(), torch.ones(C).double(), mean_ref, var_ref, rows)
    assert rel_l2(mm, emm) < 1e-5 and rel_l2(mv, emv) < 1e-5

    nb = ops.bn_bwd_nparts(xd, C, P)
    bpart = torch.zeros(nb, 2, C, dtype=torch.float64, device=dev)
    bsums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    ops.bn_bwd_reduce(dyd, xd, C, mean, rstd, scale, shift, A, P, idx, bpart)
    ops.bn_reduce_partials(bpart, bsums)
    dx = torch.zeros(B, D, D, D, C, dtype=torch.bfloat16, device=dev)
    ops.bn_bwd_apply(dyd, xd, C, mean, rstd, scale, shift, A, P, idx, bsums, float(rows), dx, pre_relu=(order == "unet"),
                     tap_other=tap_other.to(dev) if tap_other is not None else None, tap_coef=tap_coef)
    torch.cuda.synchronize()
    assert rel_l2(dx.float(), dx_ref) < 1e-2


def test_bn_fp32_c4():
    """decoder_output -> BatchNormalization -> ReLU on the fp32 4-channel tensor (lattice_vae.py:219-226)."""
    from icsg3d_b200 import ops
    B, D, C = 2, 8, 4
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, D, D, D, C, generator=g) * 2 + 1
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    dy = torch.randn(B, D, D, D, C, generator=g)
    y_ref, dx_ref, _, _ = _oracle(x, gamma, beta, "relu", "none", "vae", dy)
    dev = "cuda"
    xd, dyd = x.to(dev), dy.to(dev)
    rows = B * D ** 3
    n = ops.bn_nparts(rows, C, torch.float32)
    part = torch.zeros(n, 2, C, dtype=torch.float64, device=dev)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    mean, rstd, scale, shift = (torch.zeros(C, device=dev) for _ in range(4))
    ops.bn_stats(xd, C, part)
    ops.bn_reduce_partials(part, sums)
    ops.bn_finalize(sums, float(rows), gamma.to(dev), beta.to(dev), mean, rstd, scale, shift)
    y16 = torch.zeros(B, D, D, D, 16, dtype=torch.bfloat16, device=dev)
    y32 = torch.zeros(B, D, D, D, 4, device=dev)
    ops.bn_apply_fwd(xd, C, scale, shift, ops.ACT_RELU, ops.POST_NONE, y=y16, y32=y32)
    nb = ops.bn_bwd_nparts(xd, C, ops.POST_NONE)
    bpart = torch.zeros(nb, 2, C, dtype=torch.float64, device=dev)
    bsums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    ops.bn_bwd_reduce(dyd, xd, C, mean, rstd, scale, shift, ops.ACT_RELU, ops.POST_NONE, None, bpart)
    ops.bn_reduce_partials(bpart, bsums)
    dx = torch.zeros(B, D, D, D, 16, dtype=torch.bfloat16, device=dev)
    ops.bn_bwd_apply(dyd, xd, C, mean, rstd, scale, shift, ops.ACT_RELU, ops.POST_NONE, None, bsums, float(rows), dx)
    torch.cuda.synchronize()
    assert rel_l2(y32, y_ref) < 1e-5
    assert rel_l2(y16[..., :4].float(), y_ref) < 1e-2 and float(y16[..., 4:].abs().max()) == 0.0
    assert rel_l2(dx[..., :4].float(), dx_ref) < 1e-2
