"""Structural pins + gradient checks of the network oracle (oracle/nets.py, oracle/keras_ops.py).
The reference ships no tests or golden vectors for the networks: the oracle is pinned by what the
reference text fixes (parameter counts, shapes, loss constants) and by fp64 finite differences."""
import torch

from oracle import keras_ops as K, nets


def test_counts_shapes_constants():
    pv, pu = nets.init_vae_params(1), nets.init_unet_params(2)
    assert nets.count_trainable(pv) == 838832 and nets.count_trainable(pu) == 31156800
    assert (nets.ALPHA, nets.BETA, nets.LATENT, nets.NCOND) == (0.5, 3e-4, 256, 10)
    assert (K.BN_EPS, K.BN_MOMENTUM, K.LEAKY_ALPHA) == (1e-3, 0.99, 0.3)
    B = 2
    M = torch.rand(B, 32, 32, 32, 4)
    cond = torch.eye(10)[[1, 4]]
    taps = {}
    (loss, pm, mse, kl), xh = nets.vae_dfc_step(pv, pu, M, cond, torch.randn(B, 256), taps=taps)
    assert xh.shape == (B, 32, 32, 32, 4) and float(xh.min()) >= 0.0
    assert taps["enc_pool4"].shape == (B, 2, 2, 2, 128) and taps["enc_conv5"].shape == (B, 2, 2, 2, 4)
    assert taps["pm_x/c2"].shape == (B, 32, 32, 32, 64) and taps["pm_x/c10"].shape == (B, 4, 4, 4, 512)
    soft, sig = nets.unet_forward(pu, M[:1], training=False)
    assert soft.shape == (1, 32, 32, 32, 95) and sig.shape == (1, 32, 32, 32, 1)


def test_cond_tiling_matches_numpy_tile():
    """K.tile(cond.reshape(B,1,1,1,10), (32,32,32,4)) left-pads the multiples: channel 10*r + k = cond[k]."""
    import numpy as np
    cond = torch.eye(10)[[3, 7]]
    t = nets.tile_cond(cond, (4, 4, 4), reps=4)
    ref = np.tile(cond.numpy().reshape(2, 1, 1, 1, 10), (1, 4, 4, 4, 4))
    assert np.array_equal(t.numpy(), ref)


def test_maxpool_gradient_goes_to_first_maximum():
    x = torch.zeros(1, 2, 2, 2, 1, dtype=torch.float64, requires_grad=True)  # 8-way tie
    y = K.maxpool2(x)
    y.sum().backward()
    g = x.grad.flatten()
    assert float(g[0]) == 1.0 and float(g.sum()) == 1.0


def test_fp64_gradcheck_small_blocks():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 4, 4, 3, dtype=torch.float64, generator=g, requires_grad=True)
    k = torch.randn(3, 3, 3, 3, 2, dtype=torch.float64, generator=g, requires_grad=True)
    gamma = torch.rand(2, dtype=torch.float64, generator=g, requires_grad=True)
    beta = torch.randn(2, dtype=torch.float64, generator=g, requires_grad=True)

    def f(x, k, gamma, beta):
        y = K.conv3d_same(x, k)
        y, _, _ = K.batchnorm(y, gamma, beta, None, None, True)
        return K.upsample2(K.maxpool2(K.leaky_relu(y))).sum() + (y ** 2).mean()

    assert torch.autograd.gradcheck(f, (x, k, gamma, beta), eps=1e-6, atol=1e-5)


def test_keras_adam_first_step():
    """lr_t = lr*sqrt(1-b2)/(1-b1) at t=1 and p -= lr_t * m/(sqrt(v)+eps)."""
    opt = nets.KerasAdam(1e-3)
    p0 = torch.tensor([1.0, -2.0])
    p = {"w": p0.clone()}
    g = {"w": torch.tensor([0.5, -0.25])}
    out = opt.step(dict(p), g)
    lr_t = 1e-3 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    m, v = 0.1 * g["w"], 0.001 * g["w"] ** 2
    assert torch.allclose(out["w"], p0 - lr_t * m / (v.sqrt() + 1e-7))


def test_weighted_cce_reference_quirks():
    """Scalar weight 95 (unet.py:254) multiplies the plain CCE; clip has zero gradient outside [1e-7, 1-1e-7]."""
    logits = torch.zeros(1, 1, 1, 1, 95, requires_grad=True)
    s = torch.zeros(1, 1, 1, 1, dtype=torch.long)
    l = nets.weighted_cce(logits, s, 95.0)
    assert abs(float(l) - 95.0 * float(torch.log(torch.tensor(95.0)))) < 1e-3
    big = torch.zeros(1, 1, 1, 1, 95)
    big[..., 0] = 40.0
    big.requires_grad_(True)
    nets.weighted_cce(big, s, 95.0).backward()
    assert float(big.grad.abs().max()) == 0.0


def test_upsample_conv_fold_identity():
    """SURVEY H6, the identity csrc/conv3d_upfold.cu is built on: Conv3D(3, same) after UpSampling3D(2) == per output
    phase a 2x2x2 convolution of the low-resolution tensor with summed taps (8 of 27), borders included."""
    import torch
    from oracle import keras_ops as K
    g = torch.Generator().manual_seed(3)
    for shape, cin, cout in (((2, 4, 4, 4), 5, 3), ((1, 2, 6, 3), 2, 4)):
        x = torch.randn(*shape, cin, generator=g, dtype=torch.float64)
        w = torch.randn(3, 3, 3, cin, cout, generator=g, dtype=torch.float64)
        b = torch.randn(cout, generator=g, dtype=torch.float64)
        want = K.conv3d_same(K.upsample2(x), w, b)
        got = K.upsample2_conv3d_same_folded(x, w, b)
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)
