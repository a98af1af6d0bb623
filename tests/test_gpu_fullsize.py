"""Full-step parity at BASELINE.json's own shapes: the CUDA path against the oracle (oracle/nets.py,
torch CPU fp32) on identical inputs, weights and eps at

    configs[1]  VAE+DFC train step, 32^3, batch 32
    configs[0]  U-Net train step, 32^3, batch 8
    configs[3]  64^3: one VAE+DFC step and one U-Net step (small batch: the oracle has to finish in seconds)

bf16 operand mode.  Bars: every loss / metric |delta| <= 1e-3 * max(1, |value|) — total, PM, MSE AND
the raw KLD; U-Net total / soft / sig the same.  Activations and gradients are reported (written to
gpurun_out/fullsize_parity_*.json) and bounded loosely; the 1e-4 tier lives in tests/test_gpu_x3*.py."""
import json
import os

import pytest
import torch

from tests.util import cosine, rel_l2, synthetic_batch

pytestmark = pytest.mark.gpu


def _close(got, want, bar=1e-3):
    return all(abs(g - w) <= bar * max(1.0, abs(w)) for g, w in zip(got, want))


def _vae_case(B, d, tag):
    from icsg3d_b200.engine import VAEEngine
    from oracle import nets
    eng = VAEEngine(B, d=d, seed=3)
    M, cond, _ = synthetic_batch(B, d=d, seed=20 + d)
    eps = torch.randn(B, 256, generator=torch.Generator().manual_seed(7))
    eng.set_inputs(M.cuda(), cond.cuda(), eps.cuda())
    pv = {k: torch.from_numpy(v) for k, v in eng.vp.to_dict().items()}
    pu = {k: torch.from_numpy(v) for k, v in eng.pp.to_dict().items()}
    names = nets.trainable_names(pv)
    leaves = {k: pv[k].clone().requires_grad_(True) for k in names}
    p = dict(pv)
    p.update(leaves)
    taps = {}
    (loss, pm, mse, kl), _ = nets.vae_dfc_step(p, pu, M, cond, eps, training=True, taps=taps)
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[k] for k in names])))
    eng.train_step()
    torch.cuda.synchronize()
    got = eng.metrics_host()
    want = [float(loss), float(pm), float(mse), float(kl)]
    act = {"enc_conv1": rel_l2(eng.enc[0]["c"].float(), taps["enc_conv1"]), "z_mean": rel_l2(eng.mu, taps["z_mean"]),
           "z_log_var": rel_l2(eng.lv, taps["z_log_var"]), "x_hat": rel_l2(eng.xhat, taps["x_hat"])}
    for L in eng.pm:
        if L["tap"]:
            act["pm_x/" + L["name"]] = rel_l2(L["a"][0].float(), taps["pm_x/" + L["name"]])
            act["pm_xhat/" + L["name"]] = rel_l2(L["a"][1].float(), taps["pm_xhat/" + L["name"]])
    gr = {k: {"rel_l2": rel_l2(eng.vp.g[k], grads[k]), "cos": cosine(eng.vp.g[k], grads[k])} for k in names
          if k.endswith("kernel")}
    rep = {"config": tag, "B": B, "d": d, "metrics_cuda": got, "metrics_oracle": want,
           "abs_delta": [abs(a - b) for a, b in zip(got, want)], "act": act, "grad_kernels": gr}
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/fullsize_parity_vae_B{B}_d{d}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps({k: rep[k] for k in ("config", "metrics_cuda", "metrics_oracle", "abs_delta", "act")}, indent=1))
    print("worst kernel-gradient cosines", sorted(((v["cos"], k) for k, v in gr.items()))[:5])
    return got, want, act, gr


def test_vae_dfc_step_batch32_matches_oracle():
    """configs[1]: B=32 @ 32^3 (the headline benchmark's exact shape)."""
    got, want, act, gr = _vae_case(32, 32, "configs[1] VAE+DFC train step 32^3 batch 32")
    assert _close(got, want), (got, want)                 # [loss, PM, MSE, KLD] all at 1e-3 * max(1, |v|)
    assert act["enc_conv1"] < 1e-4 and act["z_mean"] < 1e-3 and act["x_hat"] < 3e-2 and act["pm_x/c2"] < 1e-2
    assert min(v["cos"] for v in gr.values()) > 0.95


def test_vae_dfc_step_64cubed_matches_oracle():
    """configs[3]: the 64^3 grid (batch 2 so that the oracle finishes in seconds)."""
    got, want, act, gr = _vae_case(2, 64, "configs[3] VAE+DFC train step 64^3 batch 2")
    assert _close(got, want), (got, want)
    assert act["enc_conv1"] < 1e-4 and act["x_hat"] < 3e-2
    assert min(v["cos"] for v in gr.values()) > 0.95


def _unet_case(B, d, tag):
    from icsg3d_b200.unet_engine import UNetEngine
    from oracle import nets
    eng = UNetEngine(B, d=d, seed=5, lr=1e-3)
    M, _, S = synthetic_batch(B, d=d, seed=30 + d)
    eng.set_inputs(M.cuda(), S.cuda())
    p = {k: torch.from_numpy(v) for k, v in eng.pp.to_dict().items()}
    names = nets.trainable_names(p)
    leaves = {k: p[k].clone().requires_grad_(True) for k in names}
    pp = dict(p)
    pp.update(leaves)
    taps = {}
    out, soft, sig = nets.unet_loss(pp, M, S.long(), training=True, weight=95.0, taps=taps)
    grads = dict(zip(names, torch.autograd.grad(out[0], [leaves[k] for k in names])))
    eng.keep_logits = True   # the fused head kernel keeps the logits in tensor memory; materialise them for the taps
    eng.train_step()
    torch.cuda.synchronize()
    got = eng.metrics_host()
    want = [float(o.detach()) for o in out]
    act = {n: rel_l2(eng.L[n]["a"].float(), taps[n]) for n in ("c1", "c2", "c10", "c18")}
    act["soft_logits"] = rel_l2(eng.logits[..., :95], taps["soft_logits"])
    act["sig_logit"] = rel_l2(eng.logits[..., 95:96], taps["sig_logit"])
    gr = {k: {"rel_l2": rel_l2(eng.pp.g[k], grads[k]), "cos": cosine(eng.pp.g[k], grads[k])} for k in names
          if k.endswith("kernel")}
    rep = {"config": tag, "B": B, "d": d, "metrics_cuda": got, "metrics_oracle": want,
           "abs_delta": [abs(a - b) for a, b in zip(got, want)], "act": act, "grad_kernels": gr}
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/fullsize_parity_unet_B{B}_d{d}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps({k: rep[k] for k in ("config", "metrics_cuda", "metrics_oracle", "abs_delta", "act")}, indent=1))
    print("worst kernel-gradient cosines", sorted(((v["cos"], k) for k, v in gr.items()))[:5])
    return got, want, act, gr


def test_unet_step_batch8_matches_oracle():
    """configs[0]: U-Net train step, B=8 @ 32^3."""
    got, want, act, gr = _unet_case(8, 32, "configs[0] U-Net train step 32^3 batch 8")
    assert _close(got[:3], want[:3]), (got, want)          # [total, weighted CCE (x95), BCE] at 1e-3 * max(1, |v|)
    assert act["c1"] < 1e-2 and act["soft_logits"] < 0.1
    assert gr["c18/kernel"]["cos"] > 0.99 and min(v["cos"] for v in gr.values()) > 0.8


def test_unet_step_64cubed_matches_oracle():
    """configs[3]: the 64^3 grid (batch 1 for the oracle's sake)."""
    got, want, act, gr = _unet_case(1, 64, "configs[3] U-Net train step 64^3 batch 1")
    assert _close(got[:3], want[:3]), (got, want)
    assert act["c1"] < 1e-2 and act["soft_logits"] < 0.1
    assert gr["c18/kernel"]["cos"] > 0.99 and min(v["cos"] for v in gr.values()) > 0.8
