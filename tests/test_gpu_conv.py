"""T1 kernel-local parity: tcgen05 Conv3D fprop / dgrad / wgrad through the C ABI against the
oracle's Keras-semantics conv (oracle/keras_ops.py::conv3d_same, torch CPU fp32) on the same bf16-rounded
operands, plus the on-device CUDA-core cross-check.  bf16 outputs: rel-L2 <= 1e-2 (north_star tolerance);
fp32 outputs only differ by accumulation order."""
import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu

SHAPES = [  # (B, D, cin, cout) — the layer shapes of SURVEY §8a at reduced batch
    (2, 32, 16, 16), (2, 32, 32, 64), (1, 32, 64, 32), (2, 16, 64, 128), (2, 8, 128, 256), (4, 4, 256, 512),
    (8, 2, 128, 16), (1, 2, 16, 16),
    # plane-streaming kd-folded kernel (conv3d_stream.cu): narrow layers at W >= 16, incl. odd batch / mid-column cuts
    (3, 32, 16, 32), (2, 16, 16, 32), (1, 16, 64, 32), (5, 16, 32, 16), (1, 32, 32, 16),
]


def _mk(B, D, cin, cout, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, D, D, D, cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(3, 3, 3, cin, cout, generator=g) / (27 * cin) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(cout, generator=g)
    return x, w, b


@pytest.mark.parametrize("B,D,cin,cout", SHAPES)
def test_fprop_matches_oracle(B, D, cin, cout):
    from icsg3d_b200 import ops
    from oracle import keras_ops as K
    x, w, b = _mk(B, D, cin, cout)
    ref = torch.relu(K.conv3d_same(x.float(), w, b))
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    wp = ops.pack_conv_w_fprop(wd)
    y = ops.conv3d_k3(xd, wp, bd, act=ops.ACT_RELU)
    y32 = ops.conv3d_k3(xd, wp, bd, act=ops.ACT_RELU, out_dtype=torch.float32)
    yr = ops.conv3d_k3(xd, wp, bd, act=ops.ACT_RELU, ref=True)
    torch.cuda.synchronize()
    assert rel_l2(y32, ref) < 2e-5
    assert rel_l2(yr, ref) < 2e-5
    assert rel_l2(y.float(), ref) < 1e-2


@pytest.mark.parametrize("B,D,cin,cout", SHAPES[:6] + SHAPES[8:])
def test_dgrad_matches_oracle(B, D, cin, cout):
    """Conv3DBackpropInput = the same kernel on mirrored/transposed weights."""
    from icsg3d_b200 import ops
    from oracle import keras_ops as K
    x, w, _ = _mk(B, D, cin, cout, seed=1)
    g = torch.Generator().manual_seed(2)
    dy = torch.randn(B, D, D, D, cout, generator=g).to(torch.bfloat16)
    xr = x.float().requires_grad_(True)
    K.conv3d_same(xr, w).backward(dy.float())
    wp = ops.pack_conv_w_dgrad(w.cuda())
    dx = ops.conv3d_k3(dy.cuda(), wp, None, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_l2(dx, xr.grad) < 2e-5


@pytest.mark.parametrize("B,D,cin,cout", [(2, 32, 16, 16), (1, 32, 32, 16), (2, 16, 64, 32), (2, 8, 128, 64),
                                          (4, 4, 64, 128), (8, 2, 128, 16), (4, 4, 16, 128), (2, 8, 256, 256),
                                          (3, 16, 16, 32), (1, 32, 64, 32), (5, 16, 32, 16), (1, 64, 16, 16)])
def test_wgrad_matches_oracle(B, D, cin, cout):
    from icsg3d_b200 import ops
    from oracle import keras_ops as K
    x, w, _ = _mk(B, D, cin, cout, seed=3)
    g = torch.Generator().manual_seed(4)
    dy = torch.randn(B, D, D, D, cout, generator=g).to(torch.bfloat16)
    wr = w.clone().requires_grad_(True)
    K.conv3d_same(x.float(), wr).backward(dy.float())
    dw = ops.conv3d_k3_wgrad(x.cuda(), dy.cuda())
    dw2 = ops.conv3d_k3_wgrad(x.cuda(), dy.cuda())
    torch.cuda.synchronize()
    assert rel_l2(dw.view(3, 3, 3, cin, cout), wr.grad) < 2e-5
    assert torch.equal(dw, dw2), "wgrad must be deterministic (fixed-order split reduction)"


def test_condition_fold_equals_tiled_concat():
    """Encoder conv1 on [M | one-hot tiled 4x] (lattice_vae.py:167-173, Cin=44) == 16-channel conv on
    [M | one-hot | 0 0] with the 4 replicas folded into the weights, borders included."""
    from icsg3d_b200 import ops
    from oracle import keras_ops as K, nets
    B, D = 2, 16
    g = torch.Generator().manual_seed(5)
    M = torch.randn(B, D, D, D, 4, generator=g)
    cond = torch.eye(10)[torch.tensor([3, 7])]
    w = (torch.randn(3, 3, 3, 44, 16, generator=g) / (27 * 44) ** 0.5)
    full = torch.cat([M, nets.tile_cond(cond, (D, D, D), reps=4)], dim=-1)
    ref = K.conv3d_same(full.to(torch.bfloat16).float(), w.to(torch.bfloat16).float())
    xe = torch.zeros(B, D, D, D, 16, dtype=torch.bfloat16, device="cuda")
    ops.pack_vae_input(M.cuda(), cond.cuda(), xe, None)
    wp = ops.pack_conv_w_fprop(w.cuda(), cin_pad=16, cin_lead=4, fold=4, fold_c=10)
    y = ops.conv3d_k3(xe, wp, None, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_l2(y, ref) < 5e-3  # folded weights are rounded to bf16 after the sum of 4 replicas
    # wgrad un-fold: gradient of the 4 replicas is identical
    dy = torch.randn(B, D, D, D, 16, generator=g).to(torch.bfloat16)
    dwp = ops.conv3d_k3_wgrad(xe, dy.cuda())
    dw = ops.unpack_conv_dw(dwp, 44, 16, cin_lead=4, fold=4, fold_c=10)
    fr = full.to(torch.bfloat16).float()
    wr = w.clone().requires_grad_(True)
    K.conv3d_same(fr, wr).backward(dy.float())
    torch.cuda.synchronize()
    assert rel_l2(dw, wr.grad) < 2e-5


@pytest.mark.parametrize("B,D,cin,cout", [(32, 32, 32, 64), (32, 32, 64, 32), (32, 32, 16, 16), (32, 16, 64, 32),
                                          (32, 16, 64, 64), (2, 64, 32, 64), (1, 64, 16, 32), (16, 64, 16, 16)])
def test_stream_kernel_full_batch_matches_cuda_core_reference(B, D, cin, cout):
    """Full BASELINE batch: the tcgen05 path (plane-streaming kernel for these shapes) against the CUDA-core
    cross-check kernel on the device (fp32 outputs; same operands, different accumulation order)."""
    from icsg3d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, D, D, D, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(3, 3, 3, cin, cout, device="cuda", generator=g) / (27 * cin) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    wp = ops.pack_conv_w_fprop(w)
    y = ops.conv3d_k3(x, wp, b, act=ops.ACT_LEAKY, out_dtype=torch.float32)
    yr = ops.conv3d_k3(x, wp, b, act=ops.ACT_LEAKY, ref=True)
    torch.cuda.synchronize()
    assert rel_l2(y, yr) < 2e-5
    assert float((y - yr).abs().max()) < 1e-3


@pytest.mark.parametrize("B,D,cin,cout", [(32, 32, 32, 16), (32, 32, 16, 16), (32, 16, 64, 32), (2, 64, 32, 16), (1, 64, 16, 32)])
def test_wgrad_stream_kernel_full_batch_matches_per_tap_kernel(B, D, cin, cout):
    """Full BASELINE batch: the plane-streaming filter-gradient kernel (kh folded into M, kw into N) against the
    per-tap im2col kernel (independent operand path); both accumulate in fp32 on the tensor cores."""
    from icsg3d_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(B, D, D, D, cin, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(B, D, D, D, cout, device="cuda", generator=g).to(torch.bfloat16)
    try:
        _lib.call("icsg3d_conv3d_set_impl", 1)
        ref = ops.conv3d_k3_wgrad(x, dy).clone()
    finally:
        _lib.call("icsg3d_conv3d_set_impl", 0)
    got = ops.conv3d_k3_wgrad(x, dy)
    got2 = ops.conv3d_k3_wgrad(x, dy)
    torch.cuda.synchronize()
    assert rel_l2(got, ref) < 1e-5
    assert torch.equal(got, got2)


@pytest.mark.parametrize("B,D,cin,cout,act,odt", [
    (3, 32, 32, 64, 1, "bf16"), (2, 32, 16, 16, 0, "bf16"), (5, 16, 64, 32, 2, "bf16"), (2, 32, 16, 32, 1, "bf16"),   # streaming kernel
    (2, 32, 48, 16, 0, "f32"),                                                # split-operand encoder layer, fp32 output
    (4, 16, 64, 128, 1, "bf16"), (6, 8, 128, 128, 1, "bf16"), (32, 8, 128, 256, 1, "bf16"), (32, 8, 64, 128, 2, "bf16"),  # halo kernel
    (32, 16, 64, 128, 0, "f32"), (3, 16, 64, 128, 2, "bf16")])
def test_fused_conv_bn_statistics_match_stored_output(B, D, cin, cout, act, odt):
    """conv epilogue statistics (icsg3d_conv3d_k3_igemm_stats; plane-streaming AND halo kernel) == per-channel sum / sum of
    squares of the STORED output (bf16-rounded or fp32), i.e. what a separate icsg3d_bn_stats pass over it measures."""
    from icsg3d_b200 import _lib, ops
    x, w, b = _mk(B, D, cin, cout, seed=21)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    wp = ops.pack_conv_w_fprop(wd)
    _lib.call("icsg3d_conv3d_set_halo_stats", 1)   # opt-in for the halo kernel (order-dependent fp32 atomics)
    try:
        n = ops.conv3d_k3_stats_parts(xd, wp)
        assert n > 0, "this shape must be served by a kernel with fused statistics (plane-streaming or halo)"
        dt = torch.bfloat16 if odt == "bf16" else torch.float32
        part = torch.full((n, 2, cout), float("nan"), dtype=torch.float64, device="cuda")
        y = ops.conv3d_k3(xd, wp, bd, act=act, stats=part, out_dtype=dt)
        y_plain = ops.conv3d_k3(xd, wp, bd, act=act, out_dtype=dt)
        torch.cuda.synchronize()
    finally:
        _lib.call("icsg3d_conv3d_set_halo_stats", 0)
    assert torch.equal(y, y_plain)
    yd = y.double().view(-1, cout)
    got = part.sum(0)
    assert torch.allclose(got[0], yd.sum(0), rtol=1e-5, atol=1e-3 * yd.abs().sum(0).max().item() * 1e-3)
    assert torch.allclose(got[1], (yd * yd).sum(0), rtol=1e-5)


@pytest.mark.parametrize("B,D,cin,cout,act", [(4, 4, 256, 512, 1), (8, 2, 128, 16, 2), (32, 4, 512, 512, 1), (32, 4, 64, 128, 0),
                                              (3, 4, 16, 128, 0)])
def test_split_k_per_tap_kernel_matches_unsplit(B, D, cin, cout, act):
    """icsg3d_conv3d_k3_igemm_ws: the 4^3 / 2^3 layers with the K range (taps x channel chunks) split over several CTAs
    and a fixed-order fp32 reduction == the unsplit kernel (fp32 output: accumulation order only) and deterministic."""
    from icsg3d_b200 import ops
    x, w, b = _mk(B, D, cin, cout, seed=31)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    wp = ops.pack_conv_w_fprop(wd)
    need = ops.conv3d_k3_workspace_bytes(B, D, cin, cout)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    y0 = ops.conv3d_k3(xd, wp, bd, act=act, out_dtype=torch.float32)
    y1 = ops.conv3d_k3(xd, wp, bd, act=act, out_dtype=torch.float32, ws=ws)
    y2 = ops.conv3d_k3(xd, wp, bd, act=act, out_dtype=torch.float32, ws=ws)
    yb = ops.conv3d_k3(xd, wp, bd, act=act, ws=ws)
    torch.cuda.synchronize()
    assert rel_l2(y1, y0) < 2e-5
    assert torch.equal(y1, y2)
    assert rel_l2(yb.float(), y0) < 1e-2
    if (B, D) == (32, 4):
        assert need > 0, "the 4^3 layers at batch 32 must be split"


@pytest.mark.parametrize("B,D,cin,cout,act,kind", [
    (2, 32, 16, 32, 1, 2), (3, 32, 32, 64, 1, 2),          # plane-streaming kernel
    (4, 16, 64, 128, 1, 1), (6, 8, 128, 128, 2, 1),        # halo kernel
    (8, 16, 128, 128, 1, 0), (2, 32, 192, 128, 1, 0),      # per-tap kernel (unsplit)
    (4, 4, 256, 512, 1, 0), (3, 4, 16, 128, 0, 0)])        # per-tap kernel, K split over the workspace
def test_conv_epilogue_post_affine_matches_conv_then_affine(B, D, cin, cout, act, kind):
    """icsg3d_conv3d_k3_igemm_post (inference Conv3D + ReLU + BatchNorm in one kernel, unet.py:277-279):
    y = scale * act(conv + bias) + shift in the epilogue of every conv kernel == the fp32 conv output put through the
    same affine, rounded to bf16 once; also into a channel slice of a wider buffer (the U-Net's concatenation buffers)."""
    import ctypes
    from icsg3d_b200 import _lib, ops
    x, w, b = _mk(B, D, cin, cout, seed=41)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    wp = ops.pack_conv_w_fprop(wd)
    plan = (ctypes.c_int * 10)()
    _lib.lib().icsg3d_conv3d_k3_plan(B, D, D, D, cin, cout, 148, plan)
    assert plan[0] == kind, f"shape is served by kernel {plan[0]}, the case was written for {kind}"
    g = torch.Generator(device="cuda").manual_seed(5)
    scale = torch.rand(cout, device="cuda", generator=g) * 2 - 0.5
    shift = torch.randn(cout, device="cuda", generator=g)
    need = ops.conv3d_k3_workspace_bytes(B, D, cin, cout)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    a32 = ops.conv3d_k3(xd, wp, bd, act=act, out_dtype=torch.float32, ws=ws)
    want = torch.addcmul(shift, a32, scale)
    buf = torch.full((B, D, D, D, cout + 32), 7.0, dtype=torch.bfloat16, device="cuda")
    got = buf[..., 16:16 + cout]
    ops.conv3d_k3(xd, wp, bd, act=act, out=got, post=(scale, shift), ws=ws)
    torch.cuda.synchronize()
    assert torch.all(buf[..., :16] == 7.0) and torch.all(buf[..., 16 + cout:] == 7.0)
    err = (got.float() - want).abs()
    assert float((err / (want.abs() + 1e-2)).max()) < 8e-3          # one bf16 rounding of the fp32 result
    assert rel_l2(got.float(), want) < 3e-3


@pytest.mark.parametrize("B,D,cs,cu,cout,exact", [(2, 8, 64, 128, 128, True), (1, 16, 64, 128, 128, True), (3, 8, 128, 256, 256, False),
                                                  (2, 8, 256, 512, 512, False), (5, 4, 64, 64, 64, True),
                                                  (40, 16, 64, 128, 128, True)])   # enough tiles for the paired form
def test_upsample_conv_fold_matches_materialised_upsample_concat_conv(B, D, cs, cu, cout, exact):
    """csrc/conv3d_upfold.cu (Conv3D over concatenate([skip, UpSampling3D(2)(low)]) as 27 skip taps + 8 folded
    taps per output phase) == the conv kernel on the materialised K.upsample x2 + concat tensor (unet.py:309-332)."""
    from icsg3d_b200 import ops
    g = torch.Generator().manual_seed(51)
    cin = cs + cu
    skip = torch.randn(B, D, D, D, cs, generator=g).to(torch.bfloat16).cuda()
    low = torch.randn(B, D // 2, D // 2, D // 2, cu, generator=g).to(torch.bfloat16).cuda()
    if exact:   # multiples of 1/8: the folded sums of up to 8 weights are exact in bf16 -> same products, same result
        w = (torch.randint(-4, 5, (3, 3, 3, cin, cout), generator=g).float() / 8).cuda()
    else:
        w = (torch.randn(3, 3, 3, cin, cout, generator=g) / (27 * cin) ** 0.5).cuda()
    b = (torch.randn(cout, generator=g) * 0.1).cuda()
    up = low.repeat_interleave(2, 1).repeat_interleave(2, 2).repeat_interleave(2, 3)
    cat = torch.cat([skip, up], dim=-1).contiguous()
    want32 = ops.conv3d_k3(cat, ops.pack_conv_w_fprop(w), b, act=1, out_dtype=torch.float32)
    wf = ops.pack_conv_w_upfold(w, 0, cs, cs, cu)
    got = ops.conv3d_k3_upfold(skip, low, wf, b, cout, act=1)
    # ... and reading both operands as channel slices of wider buffers, writing into a slice, with the inference affine
    sbuf = torch.zeros(B, D, D, D, cs + 64, dtype=torch.bfloat16, device="cuda")
    lbuf = torch.zeros(B, D // 2, D // 2, D // 2, cu + 64, dtype=torch.bfloat16, device="cuda")
    sbuf[..., 64:] = skip
    lbuf[..., :cu] = low
    ybuf = torch.full((B, D, D, D, cout + 16), 3.0, dtype=torch.bfloat16, device="cuda")
    scale, shift = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
    ops.conv3d_k3_upfold(sbuf[..., 64:], lbuf[..., :cu], wf, b, cout, act=1, out=ybuf[..., :cout], post=(scale, shift))
    torch.cuda.synchronize()
    tol = 1e-3 if exact else 4e-3
    assert rel_l2(got.float(), want32) < tol + 2e-3          # bf16 rounding of the output
    if exact:
        assert float((got.float() - want32).abs().max()) <= float(want32.abs().max()) * 2 ** -8
    assert rel_l2(ybuf[..., :cout].float(), torch.addcmul(shift, want32, scale)) < tol + 2e-3
    assert torch.all(ybuf[..., cout:] == 3.0)


@pytest.mark.parametrize("B,D,cs,cu,cout", [(2, 8, 64, 128, 128), (1, 16, 64, 128, 128), (3, 8, 128, 256, 256), (2, 8, 256, 512, 512),
                                            (5, 4, 64, 64, 64), (20, 32, 64, 128, 128)])   # last: paired low-resolution tiles
def test_upsample_conv_fold_dgrad_low_matches_upsample_backward_of_dgrad(B, D, cs, cu, cout):
    """Backward of the folded conv w.r.t. its low-resolution input == UpSampling3D's backward (sum over the 8 children) of
    the data-gradient conv on the materialised concat tensor (weights exactly representable: same products)."""
    from icsg3d_b200 import ops
    g = torch.Generator().manual_seed(61)
    cin = cs + cu
    dy = torch.randn(B, D, D, D, cout, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randint(-4, 5, (3, 3, 3, cin, cout), generator=g).float() / 8).cuda()
    dcat32 = ops.conv3d_k3(dy, ops.pack_conv_w_dgrad(w), None, out_dtype=torch.float32)           # [B,D,D,D,cin]
    want = dcat32[..., cs:].reshape(B, D // 2, 2, D // 2, 2, D // 2, 2, cu).sum(dim=(2, 4, 6))
    wt = ops.pack_conv_w_upfold_dgrad(w, cs, cu)
    buf = torch.full((B, D // 2, D // 2, D // 2, cu + 32), 5.0, dtype=torch.bfloat16, device="cuda")
    ops.conv3d_k3_upfold_dgrad_low(dy, wt, cu, out=buf[..., 16:16 + cu])
    torch.cuda.synchronize()
    got = buf[..., 16:16 + cu].float()
    assert rel_l2(got, want) < 3e-3
    assert float((got - want).abs().max()) <= float(want.abs().max()) * 2 ** -7
    assert torch.all(buf[..., :16] == 5.0) and torch.all(buf[..., 16 + cu:] == 5.0)


@pytest.mark.parametrize("B,D,cin,cout,act", [(3, 32, 128, 128, 1), (5, 32, 64, 128, 0), (40, 16, 192, 128, 2), (20, 16, 128, 64, 1)])
def test_per_tap_kernel_with_paired_m_tiles_matches_cuda_core_reference(B, D, cin, cout, act):
    """Per-tap kernel in its paired form (two 128-voxel tiles share every weight tile, conv3d_igemm.cu: mt = 2; taken when
    the layer has >= 2 x SMs tile pairs) against the CUDA-core cross-check kernel, and bit-equal to the unpaired form."""
    import ctypes
    from icsg3d_b200 import _lib, ops
    x, w, b = _mk(B, D, cin, cout, seed=71)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    wp = ops.pack_conv_w_fprop(wd)
    _lib.call("icsg3d_conv3d_set_impl", 1)          # per-tap kernel regardless of the dispatch rules
    try:
        got = ops.conv3d_k3(xd, wp, bd, act=act, out_dtype=torch.float32)
        gb = ops.conv3d_k3(xd, wp, bd, act=act)
    finally:
        _lib.call("icsg3d_conv3d_set_impl", 0)
    ref = ops.conv3d_k3(xd, wp, bd, act=act, ref=True)
    torch.cuda.synchronize()
    assert rel_l2(got, ref) < 1e-5
    assert rel_l2(gb.float(), ref) < 4e-3

