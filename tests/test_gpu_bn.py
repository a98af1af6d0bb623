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
    if tap_other is not None:
        # fused DFC feature loss: the apply pass also returns sum (x - tap_other)^2 as per-block fp64 partials, dx unchanged
        nsq = ops.bn_bwd_apply_nblocks(xd, C, P)
        sq = torch.full((nsq,), -1.0, dtype=torch.float64, device=dev)
        dx_sq = torch.zeros_like(dx)
        ops.bn_bwd_apply(dyd, xd, C, mean, rstd, scale, shift, A, P, idx, bsums, float(rows), dx_sq, pre_relu=True,
                         tap_other=tap_other.to(dev), tap_coef=tap_coef, tap_sq=sq)
        torch.cuda.synchronize()
        want_sq = ((x.double() - tap_other.double()) ** 2).sum().item()
        assert abs(sq.sum().item() - want_sq) <= 1e-5 * want_sq, (sq.sum().item(), want_sq)
        assert torch.equal(dx_sq, dx)
    # the same backward in ONE cooperative launch (reduce -> grid barrier -> fixed-order sums -> grid barrier -> apply)
    nf = ops.bn_bwd_fused_nparts(C, torch.bfloat16)
    fpart = torch.zeros(nf, 2, C, dtype=torch.float64, device=dev)
    fsums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dx2 = torch.zeros_like(dx)
    ops.bn_bwd_fused(dyd, xd, C, mean, rstd, scale, shift, A, P, idx, fpart, fsums, float(rows), dx2, dgamma=dg, dbeta=db,
                     pre_relu=(order == "unet"), tap_other=tap_other.to(dev) if tap_other is not None else None,
                     tap_coef=tap_coef)
    torch.cuda.synchronize()
    assert rel_l2(dx2.float(), dx_ref) < 1e-2
    assert torch.allclose(fsums, bsums, rtol=1e-9, atol=1e-9)
    assert rel_l2(dx2.float(), dx.float()) < 1e-3
    assert torch.allclose(db.double(), bsums[:C], rtol=1e-5, atol=1e-5) and torch.allclose(dg.double(), bsums[C:], rtol=1e-5, atol=1e-5)


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
    nf = ops.bn_bwd_fused_nparts(C, torch.float32)
    fpart = torch.zeros(nf, 2, C, dtype=torch.float64, device=dev)
    fsums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    dx2 = torch.zeros_like(dx)
    ops.bn_bwd_fused(dyd, xd, C, mean, rstd, scale, shift, ops.ACT_RELU, ops.POST_NONE, None, fpart, fsums, float(rows), dx2)
    torch.cuda.synchronize()
    assert rel_l2(dx2[..., :4].float(), dx_ref) < 1e-2 and torch.allclose(fsums, bsums, rtol=1e-9, atol=1e-9)


def test_peer_memory_allreduce_kernel_two_emulated_ranks():
    """csrc/bn.cu::bn_reduce_allreduce_kernel (sync-BN statistics over NVLink peer memory) with two "ranks"
    emulated on ONE device: two buffers in the same address space, the two launches on different streams (they spin on
    each other's flags, so they must run concurrently).  Checks the protocol (push, flags, parity double buffering across
    epochs), that both ranks get bit-identical global sums, and the finalisation against torch."""
    from icsg3d_b200 import ops
    dev = torch.device("cuda")
    world, nslots, cmax, C = 2, 4, 128, 64
    nbytes = ops.bn_allreduce_buffer_bytes(world, nslots, cmax)
    bufs = [torch.zeros(nbytes // 8, dtype=torch.int64, device=dev) for _ in range(world)]
    peers = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    epoch = torch.zeros(1, dtype=torch.int64, device=dev)
    streams = [torch.cuda.Stream() for _ in range(world)]
    g = torch.Generator(device="cuda").manual_seed(3)
    for step in range(3):
        epoch.add_(1)
        parts = [torch.randn(37 + 5 * r, 2, C, dtype=torch.float64, device=dev, generator=g).abs() * 100 for r in range(world)]
        outs = []
        torch.cuda.synchronize()
        for r in range(world):
            o = dict(sums=torch.zeros(2 * C, dtype=torch.float64, device=dev),
                     **{k: torch.zeros(C, device=dev) for k in ("mean", "rstd", "scale", "shift")})
            with torch.cuda.stream(streams[r]):
                ops.bn_reduce_allreduce_finalize(parts[r], 1000.0, None, None, o["sums"], o["mean"], o["rstd"], o["scale"],
                                                 o["shift"], peers, world, r, slot=step % nslots, nslots=nslots, cmax=cmax,
                                                 epoch=epoch)
            outs.append(o)
        torch.cuda.synchronize()
        want = sum(p.sum(0) for p in parts).reshape(-1)
        assert torch.equal(outs[0]["sums"], outs[1]["sums"]), "ranks must agree bit for bit"
        assert torch.allclose(outs[0]["sums"], want, rtol=1e-12)
        mean = want[:C] / 1000.0
        var = (want[C:] / 1000.0 - mean * mean).clamp_min(0)
        assert torch.allclose(outs[1]["mean"].double(), mean, rtol=1e-6)
        assert torch.allclose(outs[1]["rstd"].double(), 1.0 / torch.sqrt(var + 1e-3), rtol=1e-6)
    # backward flavour: global sums + LOCAL dgamma/dbeta
    epoch.add_(1)
    parts = [torch.randn(19, 2, C, dtype=torch.float64, device=dev, generator=g) for _ in range(world)]
    res = []
    for r in range(world):
        sg = torch.zeros(2 * C, dtype=torch.float64, device=dev)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        with torch.cuda.stream(streams[r]):
            ops.bn_reduce_allreduce_grads(parts[r], sg, peers, world, r, slot=3, nslots=nslots, cmax=cmax, epoch=epoch,
                                          dgamma=dg, dbeta=db)
        res.append((sg, dg, db))
    torch.cuda.synchronize()
    want = sum(p.sum(0) for p in parts).reshape(-1)
    assert torch.equal(res[0][0], res[1][0]) and torch.allclose(res[0][0], want, rtol=1e-12, atol=1e-12)
    for r in range(world):
        loc = parts[r].sum(0)
        assert torch.allclose(res[r][2].double(), loc[0], rtol=1e-6, atol=1e-6)  # dbeta = local sum g
        assert torch.allclose(res[r][1].double(), loc[1], rtol=1e-6, atol=1e-6)  # dgamma = local sum g*xhat


def test_peer_memory_gradient_allreduce_adam_two_emulated_ranks():
    """csrc/misc.cu::adam_allreduce_kernel: gradient all-reduce over peer memory fused with Keras-Adam, two ranks emulated
    on one device (two streams).  Both ranks must end with bit-identical parameters == plain adam_keras_step on g0+g1."""
    from icsg3d_b200 import ops
    dev = torch.device("cuda")
    world, n = 2, 3 * 4096 + 1234
    gbytes = ops.adam_allreduce_buffer_bytes(world, n)
    bufs = [torch.zeros((gbytes + 7) // 8, dtype=torch.int64, device=dev) for _ in range(world)]
    peers = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    epoch = torch.zeros(1, dtype=torch.int64, device=dev)
    streams = [torch.cuda.Stream() for _ in range(world)]
    gen = torch.Generator(device="cuda").manual_seed(9)
    p0 = torch.randn(n, device=dev, generator=gen)
    st = [dict(p=p0.clone(), m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev),
               state=torch.zeros(2, dtype=torch.float64, device=dev)) for _ in range(world + 1)]
    for step in range(3):
        epoch.add_(1)
        gs = [torch.randn(n, device=dev, generator=gen) * 1e-2 for _ in range(world)]
        gsum = gs[0] + gs[1]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ops.adam_keras_allreduce_step(st[r]["p"], gs[r], st[r]["m"], st[r]["v"], st[r]["state"], 5e-4, peers, world, r,
                                              epoch)
        ref = st[world]
        ops.adam_keras_step(ref["p"], gsum.clone(), ref["m"], ref["v"], ref["state"], 5e-4)
        torch.cuda.synchronize()
        assert torch.equal(gs[0], gsum) and torch.equal(gs[1], gsum), "g must hold the global sum on every rank"
        assert torch.equal(st[0]["p"], st[1]["p"]) and torch.equal(st[0]["v"], st[1]["v"]), "ranks must agree bit for bit"
        # vs the single-process kernel: same formula, separately compiled (FMA contraction may differ in the last bit)
        assert torch.allclose(st[0]["p"], ref["p"], rtol=1e-6, atol=1e-7), float((st[0]["p"] - ref["p"]).abs().max())
        assert torch.allclose(st[0]["v"], ref["v"], rtol=1e-5, atol=1e-12)
