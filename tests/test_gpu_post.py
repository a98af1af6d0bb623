"""SURVEY §8f.1 / §8f.3 and the metric closures of §8b on the GPU, through the C ABI, against golden vectors produced by the
reference's own functions (tests/golden/make_post_golden.py: utils.to_lattice_params, to_voxel_params,
random_rotation_3d) and against numpy restatements of unet.py:159-193."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _golden():
    from tests.golden.make_post_golden import lattice_inputs
    z = np.load(os.path.join(GOLD, "voxel_golden.npz"))
    g = np.load(os.path.join(GOLD, "post_golden.npz"))
    return z, g, lattice_inputs(z)


def test_lattice_and_voxel_params_bit_exact_vs_reference():
    """min/max are exact, the arithmetic replays numpy's float32 / float64 ops one by one: bit-identical lp and dv."""
    from icsg3d_b200 import utils
    z, g, inputs = _golden()
    for name, p in inputs.items():
        lp = utils.to_lattice_params(p)                       # numpy in -> numpy out (reference signature)
        assert lp.dtype == g[f"lat_{name}_lp"].dtype and lp.shape == g[f"lat_{name}_lp"].shape
        assert np.array_equal(lp, g[f"lat_{name}_lp"]), (name, lp, g[f"lat_{name}_lp"])
        dv = utils.to_voxel_params(lp)
        assert np.array_equal(dv, g[f"lat_{name}_dv"])
        # device form: lp and dv from the same pair of launches; the quirk a*(1-1/d): a 4 A cell reads 3.875
        lpd, dvd = utils.lattice_params_device(torch.from_numpy(np.ascontiguousarray(p)).cuda())
        assert np.array_equal(lpd.cpu().numpy(), g[f"lat_{name}_lp"]) and np.array_equal(dvd.cpu().numpy(), g[f"lat_{name}_dv"])


def test_fused_unet_input_pack_and_lattice_params():
    """The fused pass (decoder output read once): min/max of channels 1..3 of an fp32 (B,d,d,d,4) tensor + bf16 pack."""
    from icsg3d_b200 import ops, utils
    z, g, inputs = _golden()
    p = inputs["noisy32"]
    B, d = p.shape[0], p.shape[1]
    dens = np.random.default_rng(0).random((B, d, d, d, 1)).astype(np.float32)
    x = torch.from_numpy(np.concatenate([dens, p], axis=-1)).cuda()
    x16 = torch.full((B, d, d, d, 16), 7.0, dtype=torch.bfloat16, device="cuda")
    lp, dv = utils.lattice_params_device(x, c0=1, x16=x16)
    assert np.array_equal(lp.cpu().numpy(), g["lat_noisy32_lp"]) and np.array_equal(dv.cpu().numpy(), g["lat_noisy32_dv"])
    want = torch.zeros_like(x16)
    ops.pack_vae_input(x, None, None, want)
    assert torch.equal(x16, want)


def test_rotation_matches_reference_random_rotation_3d():
    from icsg3d_b200 import utils
    z, g, _ = _golden()
    k = 0
    while f"rot{k}_seq" in g.files:
        M, S, p = z[f"c{k}_M"], z[f"c{k}_S"].astype(np.float64), z[f"c{k}_p"]
        np.random.seed(100 + k)
        Mr, Sr, pr = utils.random_rotation_3d(M, S, p)      # drop-in signature; consumes np.random like the reference
        assert Sr.dtype == np.float64 and np.array_equal(Sr, g[f"rot{k}_S"].astype(np.float64)), k     # bit-exact species
        if f"rot{k}_M" in g.files:
            assert np.abs(Mr - g[f"rot{k}_M"]).max() <= 1e-12 * max(1.0, np.abs(M).max())   # spline round-off of the reference
            assert np.abs(pr - g[f"rot{k}_p"]).max() <= 1e-12 * max(1.0, np.abs(p).max())
        k += 1
    assert k >= 9


def test_rotation_batch_payloads_and_group_properties():
    """Per-sample transforms on the fused network payloads: fp32 x 4 channels (16 B/voxel), uint8 species, fp64; four
    quarter turns about one axis are the identity; a rotation is a bijection (sorted values unchanged)."""
    from icsg3d_b200 import utils
    g = torch.Generator().manual_seed(0)
    B, d = 5, 32
    x = torch.randn(B, d, d, d, 4, generator=g).cuda()
    s = torch.randint(0, 95, (B, d, d, d), generator=g, dtype=torch.uint8).cuda()
    seqs = [[utils.ROT_AXES[i] for i in np.random.default_rng(b).integers(0, 3, 3)] for b in range(B)]
    xf = [utils.rot90_transform(q) for q in seqs]
    xr, sr = utils.rotate90_batch(x, xf), utils.rotate90_batch(s, xf)
    for b in range(B):
        want_x, want_s = x[b].cpu().numpy(), s[b].cpu().numpy()
        for ax in seqs[b]:
            want_x, want_s = np.rot90(want_x, 1, axes=ax), np.rot90(want_s, 1, axes=ax)
        assert np.array_equal(xr[b].cpu().numpy(), want_x) and np.array_equal(sr[b].cpu().numpy(), want_s)
    y = x.double()
    for _ in range(4):
        y = utils.rotate90_batch(y, utils.rot90_transform([(0, 2)]))
    assert torch.equal(y, x.double())
    assert torch.equal(xr.flatten(1).sort(dim=1).values, x.flatten(1).sort(dim=1).values)


def test_heads_predict_labels_and_mask():
    from icsg3d_b200 import ops
    g = torch.Generator().manual_seed(1)
    M = 4 * 16 ** 3
    logits = torch.randn(M, 96, generator=g) * 3
    logits[5, 10] = logits[5, 20] = 50.0          # exact tie: first index wins (np.argmax)
    logits[:64, 95] = torch.linspace(-4, 4, 64)
    lg = logits.cuda()
    am = torch.empty(M, dtype=torch.uint8, device="cuda")
    mk = torch.empty(M, dtype=torch.uint8, device="cuda")
    sp = torch.empty(M, dtype=torch.float32, device="cuda")
    ops.heads_predict(lg, 95, 0.8, argmax=am, mask=mk, sig_prob=sp)
    probs = torch.softmax(lg[:, :95], dim=1)      # labels = argmax of the float32 softmax output, first index on ties
    assert float((am.cpu().long() != logits[:, :95].argmax(dim=1)).float().mean()) < 1e-3 and int(am[5]) == 10
    pa = probs.gather(1, am.long()[:, None])[:, 0]
    assert torch.allclose(pa, probs.max(dim=1).values, rtol=1e-6)
    want_sig = torch.sigmoid(logits[:, 95])
    assert torch.allclose(sp.cpu(), want_sig, atol=1e-6)
    clear = (want_sig - 0.8).abs() > 1e-5
    assert torch.equal(mk.cpu().bool()[clear], (want_sig >= 0.8)[clear])
    assert torch.equal(mk.bool(), sp >= 0.8)       # mask and probability are consistent with each other

@pytest.mark.parametrize("shape,cin,classes,f16", [((3, 4, 4, 4), 128, 95, False), ((2, 16, 16, 16), 128, 95, False),
                                                   ((1, 8, 8, 8), 64, 20, False), ((1, 8, 8, 8), 384, 95, False),
                                                   ((2, 8, 8, 8), 384, 95, True)])
def test_fused_heads_predict_equals_conv_then_predict(shape, cin, classes, f16):
    """csrc/heads_fused.cu == 1x1x1 head conv (fp32 logits) + heads_predict, bit for bit (labels, mask, probability),
    including exact ties between classes and a row count that is not a multiple of the 128-row tile."""
    from icsg3d_b200 import ops
    B, D, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    nout = (classes + 1 + 15) // 16 * 16
    x = (torch.randn(B, D, H, W, cin, device="cuda", generator=g) * 2).to(torch.bfloat16)
    w_soft = torch.randn(1, 1, 1, cin, classes, device="cuda", generator=g) / cin ** 0.5 * 3
    w_soft[..., 7] = w_soft[..., 3]                       # classes 3 and 7 tie everywhere: the first index must win
    w_sig = torch.randn(1, 1, 1, cin, 1, device="cuda", generator=g) / cin ** 0.5
    b_soft = torch.randn(classes, device="cuda", generator=g) * 0.1
    b_soft[3] += 2.0                                      # ... and often enough to be seen at every test size
    b_soft[7] = b_soft[3]
    if classes > 10:                                      # class 5 a hair BELOW class 9 and before it: its probability
        w_soft[..., 5] = w_soft[..., 9]                   # rounds to the same float32 for some voxels only (the rare
        b_soft[9] += 2.0                                  # path of the fused kernel: exp() of a 1-ulp logit difference)
        b_soft[5] = b_soft[9] - 1.2e-7
    b_sig = torch.randn(1, device="cuda", generator=g)
    wf = torch.zeros(1, nout, cin, dtype=torch.bfloat16, device="cuda")
    wd = torch.zeros(1, cin, nout, dtype=torch.bfloat16, device="cuda")
    bias = torch.zeros(nout, dtype=torch.float32, device="cuda")
    ops.pack_heads_w(w_soft, w_sig, b_soft, b_sig, wf, wd, bias)
    logits = torch.empty(B, D, H, W, nout, dtype=torch.float32, device="cuda")
    if f16:  # IEEE fp16 operand bits in the same 2-byte buffers (fp32-class split mode), weights pre-scaled by 2^10
        x = x.float().half().view(torch.bfloat16)
        wf = (wf.float() * ops.SPLIT_WSCALE).half().view(torch.bfloat16)
        ops.conv3d_k3(x, wf, bias, out=logits, split=True, fmt=1)
    else:
        ops.conv3d_k3(x, wf, bias, out=logits)
    M = B * D * H * W
    ref = [torch.empty(M, dtype=dt, device="cuda") for dt in (torch.uint8, torch.uint8, torch.float32)]
    got = [torch.full((M,), 255, dtype=torch.uint8, device="cuda"), torch.full((M,), 255, dtype=torch.uint8, device="cuda"),
           torch.full((M,), -1.0, dtype=torch.float32, device="cuda")]
    ops.heads_predict(logits, classes, 0.6, argmax=ref[0], mask=ref[1], sig_prob=ref[2])
    ops.heads_predict_fused(x, wf, bias, classes, 0.6, argmax=got[0], mask=got[1], sig_prob=got[2], f16=f16,
                            out_scale=1.0 / ops.SPLIT_WSCALE if f16 else 1.0)
    torch.cuda.synchronize()
    for r, o in zip(ref, got):
        assert torch.equal(r, o)
    assert int((ref[0] == 7).sum()) == 0 and int((ref[0] == 3).sum()) > 0     # the tie really occurs and resolves to 3
    if classes > 10 and M >= 1024:
        assert int((ref[0] == 5).sum()) > 0 and int((ref[0] == 9).sum()) > 0  # near-ties resolve both ways
    lab = logits.view(M, nout)[:, :classes].argmax(dim=1)
    assert float((got[0].long() != lab).float().mean()) < 1e-3


def test_fused_heads_predict_on_a_channel_slice_with_row_stride():
    """Features living in a wider buffer (ldx > cin), as the U-Net's concatenation buffers do."""
    from icsg3d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    big = (torch.randn(2, 8, 8, 8, 192, device="cuda", generator=g)).to(torch.bfloat16)
    x = big[..., 64:192]
    wf = (torch.randn(1, 96, 128, device="cuda", generator=g) / 6).to(torch.bfloat16)
    bias = torch.randn(96, device="cuda", generator=g) * 0.1
    logits = torch.empty(2, 8, 8, 8, 96, dtype=torch.float32, device="cuda")
    ops.conv3d_k3(x, wf, bias, out=logits)
    M = 1024
    a0, a1 = torch.empty(M, dtype=torch.uint8, device="cuda"), torch.empty(M, dtype=torch.uint8, device="cuda")
    m0, m1 = torch.empty_like(a0), torch.empty_like(a0)
    ops.heads_predict(logits, 95, 0.8, argmax=a0, mask=m0)
    ops.heads_predict_fused(x, wf, bias, 95, 0.8, argmax=a1, mask=m1)
    assert torch.equal(a0, a1) and torch.equal(m0, m1)

@pytest.mark.parametrize("shape,classes", [((2, 16, 16, 16), 95), ((3, 4, 4, 4), 95), ((1, 8, 8, 8), 20)])
def test_fused_heads_loss_equals_conv_then_heads_loss(shape, classes):
    """csrc/heads_fused.cu, training form == 1x1x1 head conv (fp32 logits) + heads_loss: losses, f1 / weighted-recall
    counts, arg-max, sigmoid probability and the bf16 gradient w.r.t. the logits."""
    from icsg3d_b200 import ops
    B, D, H, W = shape
    cin, M = 128, B * D * H * W
    g = torch.Generator(device="cuda").manual_seed(13)
    nout = (classes + 1 + 15) // 16 * 16
    x = (torch.randn(B, D, H, W, cin, device="cuda", generator=g) * 2).to(torch.bfloat16)
    w_soft = torch.randn(1, 1, 1, cin, classes, device="cuda", generator=g) / cin ** 0.5 * 2
    w_sig = torch.randn(1, 1, 1, cin, 1, device="cuda", generator=g) / cin ** 0.5
    b_soft = torch.randn(classes, device="cuda", generator=g) * 0.1
    b_sig = torch.randn(1, device="cuda", generator=g)
    wf = torch.zeros(1, nout, cin, dtype=torch.bfloat16, device="cuda")
    wd = torch.zeros(1, cin, nout, dtype=torch.bfloat16, device="cuda")
    bias = torch.zeros(nout, dtype=torch.float32, device="cuda")
    ops.pack_heads_w(w_soft, w_sig, b_soft, b_sig, wf, wd, bias)
    species = torch.randint(0, classes, (M,), device="cuda", generator=g).to(torch.uint8)
    species[::3] = 0
    cw = torch.rand(classes, device="cuda", generator=g) * 3 + 0.5
    logits = torch.empty(B, D, H, W, nout, dtype=torch.float32, device="cuda")
    ops.conv3d_k3(x, wf, bias, out=logits)
    p0 = torch.zeros(ops.heads_loss_nparts(M), 6, dtype=torch.float64, device="cuda")
    a0, s0 = torch.empty(M, dtype=torch.uint8, device="cuda"), torch.empty(M, dtype=torch.float32, device="cuda")
    d0 = torch.zeros(M, nout, dtype=torch.bfloat16, device="cuda")
    ops.heads_loss(logits, classes, species, cw, 1.0 / M, p0, argmax_out=a0, sig_prob=s0, dlogits=d0)
    p1 = torch.zeros(ops.heads_loss_fused_nparts(M), 6, dtype=torch.float64, device="cuda")
    a1, s1 = torch.full_like(a0, 255), torch.full_like(s0, -1.0)
    d1 = torch.full((M, nout), 9.0, dtype=torch.bfloat16, device="cuda")
    ops.heads_loss_fused(x, wf, bias, classes, species, cw, 1.0 / M, p1, argmax_out=a1, sig_prob=s1, dlogits=d1)
    m0, m1 = torch.zeros(5, device="cuda"), torch.zeros(5, device="cuda")
    r0, r1 = torch.zeros(6, dtype=torch.float64, device="cuda"), torch.zeros(6, dtype=torch.float64, device="cuda")
    ops.heads_loss_finalize(p0, float(M), m0, r0)
    ops.heads_loss_finalize(p1, float(M), m1, r1)
    torch.cuda.synchronize()
    assert torch.equal(a0, a1) and torch.equal(s0, s1)
    assert torch.equal(r0[2:], r1[2:])                                   # the four metric counts are integers
    assert torch.allclose(r0[:2], r1[:2], rtol=1e-6) and torch.allclose(m0, m1, rtol=1e-5)
    diff = (d0.float() - d1.float()).abs()
    assert float(diff.max()) <= float(d0.float().abs().max()) * 2 ** -7   # one bf16 ulp of the largest entry
    assert float((d0 != d1).float().mean()) < 0.02                        # soft-max sum order differs in the last bit only


def test_metric_functions_match_keras_formulas():
    """unet.py:159-193 on one-hot truth / softmax predictions: numpy restatement of the K.round(K.clip()) sums."""
    from icsg3d_b200.unet import unet as U
    rng = np.random.default_rng(3)
    n, C = 5000, 95
    lab = rng.integers(0, C, n)
    lab[: n // 2] = 0
    yt = np.eye(C, dtype=np.float32)[lab].reshape(5, 10, 10, 10, C)
    logits = rng.normal(0, 1, (n, C)).astype(np.float32)
    logits[np.arange(n), lab] += rng.choice([0.0, 6.0], n)
    e = np.exp(logits - logits.max(1, keepdims=True))
    yp = (e / e.sum(1, keepdims=True)).astype(np.float32).reshape(yt.shape)
    yp.reshape(-1)[:4] = [0.5, 1.5, -0.2, 0.50001]          # half-to-even, clip
    rc = lambda a: np.rint(np.clip(a, 0, 1))
    w = np.ones(C, np.float32); w[0] = 0
    tp, pos, pred = rc(yt * yp).sum(), rc(yt).sum(), rc(yp).sum()
    eps = 1e-7
    r, p = tp / (pos + eps), tp / (pred + eps)
    assert abs(U.r_m(yt, yp) - r) < 1e-9 and abs(U.p_m(yt, yp) - p) < 1e-9
    assert abs(U.f1_m(yt, yp) - 2 * (p * r) / (p + r + eps)) < 1e-9
    assert abs(U.wr_m(yt, yp) - rc(w * yt * yp).sum() / (rc(w * yt).sum() + eps)) < 1e-9


def test_generate_pipeline_device_resident_matches_stepwise_api():
    """generate.py:202-225 as one captured device-resident step == the same steps through the public per-call API
    (decoder.predict -> to_lattice_params -> predict_labels), bit for bit."""
    from icsg3d_b200 import utils
    from icsg3d_b200.pipeline import GeneratePipeline
    from icsg3d_b200.unet.unet import AtomUnet
    from icsg3d_b200.vae.lattice_vae import LatticeDFCVAE
    B = 4
    vae = LatticeDFCVAE(perceptual_model=None, seed=1)
    vae._set_model(batch_size=B)
    unet = AtomUnet(seed=2)
    rng = np.random.default_rng(0)
    z = rng.normal(0, 0.5, (B, 256)).astype(np.float32)
    cond = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
    pipe = GeneratePipeline(vae, unet, B)
    for _ in range(2):  # second run = graph replay
        r = pipe.run(z, cond)
        torch.cuda.synchronize()
    Mp = vae.decoder.predict([z, cond])
    assert np.array_equal(r["M_prime"].cpu().numpy(), Mp)
    lp = utils.to_lattice_params(Mp[..., 1:])
    assert np.array_equal(r["lattice"].cpu().numpy(), lp)
    assert np.array_equal(r["voxel"].cpu().numpy(), utils.to_voxel_params(lp))
    lab, mask = unet.predict_labels(Mp)
    assert np.array_equal(r["species"].cpu().numpy(), lab) and np.array_equal(r["mask"].cpu().numpy().astype(bool), mask)
