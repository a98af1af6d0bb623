"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement (plain PyTorch, fp32 or fp64) of the two reference graphs and their losses:

  * conditional DFC-VAE           — vae/lattice_vae.py:53-66 (sampling), :160-195 (encoder),
                                    :197-230 (decoder), :232-270 (losses)
  * 3-D U-Net with dual head      — unet/unet.py:272-355 (graph), :196-221 (weighted CCE),
                                    :159-193 (metrics)
  * Keras-form Adam               — keras.optimizers.Adam as used at lattice_vae.py:98, unet.py:245

Parity status: UNPINNED (see oracle/keras_ops.py).  Structural pins: parameter counts 838,832 (VAE) and
31,156,800 (U-Net), tap shapes, loss constants — derivable from the reference text.

Parameters are a flat dict name -> tensor in Keras layouts.  Every intermediate the parity tests look at
is returned in a `taps` dict under the names listed in SURVEY.md §8c.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import keras_ops as K

VAE_FILTERS = [16, 32, 64, 128]          # lattice_vae.py:93
LATENT = 256                             # lattice_vae.py:94
NCOND = 10                               # lattice_vae.py:101 (cond_shape)
ALPHA, BETA = 0.5, 3e-4                  # lattice_vae.py:96-97
UNET_BLOCKS = [                          # (name, Cin, Cout) unet.py:276-336; Cin of c1 is the input channel count
    ("c1", None, 32), ("c2", 32, 64), ("c3", 64, 64), ("c4", 64, 128), ("c5", 128, 128), ("c6", 128, 256),
    ("c9", 256, 512), ("c10", 512, 512), ("c13", 768, 512), ("c14", 512, 256), ("c15", 384, 256),
    ("c16", 256, 128), ("c17", 192, 128), ("c18", 128, 128),
]
PM_TAPS = ["c2", "c4", "c6", "c10"]      # re_lu_2,4,6,8 = ReLU outputs of c2,c4,c6,c10 (lattice_vae.py:100; R12)
PM_PREFIX = ["c1", "c2", "c3", "c4", "c5", "c6", "c9", "c10"]


# --------------------------------------------------------------------------------------------------
# parameter construction (Keras initialisers: glorot_uniform kernels, zero biases, BN gamma=1, beta=0,
# moving_mean=0, moving_variance=1)
# --------------------------------------------------------------------------------------------------
def _conv_params(p, name, cin, cout, gen, dtype, k=3):
    fan_in, fan_out = k ** 3 * cin, k ** 3 * cout
    p[name + "/kernel"] = K.glorot_uniform((k, k, k, cin, cout), fan_in, fan_out, gen).to(dtype)
    p[name + "/bias"] = torch.zeros(cout, dtype=dtype)


def _bn_params(p, name, c, dtype):
    p[name + "/gamma"] = torch.ones(c, dtype=dtype)
    p[name + "/beta"] = torch.zeros(c, dtype=dtype)
    p[name + "/moving_mean"] = torch.zeros(c, dtype=dtype)
    p[name + "/moving_variance"] = torch.ones(c, dtype=dtype)


def _dense_params(p, name, cin, cout, gen, dtype):
    p[name + "/kernel"] = K.glorot_uniform((cin, cout), cin, cout, gen).to(dtype)
    p[name + "/bias"] = torch.zeros(cout, dtype=dtype)


def init_vae_params(seed=1, dtype=torch.float32, channels=4, ncond=NCOND, d=32):
    """lattice_vae.py:160-230.  Encoder conv1 sees channels + 4*ncond inputs (K.tile left-pads its
    multiples, so the one-hot is repeated input_shape[-1]=4 times along C — SURVEY R1)."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    cin = channels + channels * ncond
    for i, f in enumerate(VAE_FILTERS, 1):
        _conv_params(p, f"enc_conv{i}", cin, f, gen, dtype)
        _bn_params(p, f"enc_bn{i}", f, dtype)
        cin = f
    _conv_params(p, "enc_conv5", cin, 4, gen, dtype)
    s = d // 16
    _dense_params(p, "enc_dense", s * s * s * 4, LATENT, gen, dtype)
    _dense_params(p, "z_mean", LATENT, LATENT, gen, dtype)
    _dense_params(p, "z_log_var", LATENT, LATENT, gen, dtype)
    s0 = d // 8
    _dense_params(p, "dec_dense", LATENT + ncond, s0 * s0 * s0 * 4, gen, dtype)
    cin = 4
    for i, f in enumerate(VAE_FILTERS[::-1], 1):
        _conv_params(p, f"dec_conv{i}", cin, f, gen, dtype)
        _bn_params(p, f"dec_bn{i}", f, dtype)
        cin = f
    _conv_params(p, "decoder_output", cin, channels, gen, dtype)
    _bn_params(p, "dec_bn5", channels, dtype)
    return p


def init_unet_params(seed=1, dtype=torch.float32, channels=4, classes=95):
    """unet.py:272-355."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, cin, cout in UNET_BLOCKS:
        _conv_params(p, name, channels if cin is None else cin, cout, gen, dtype)
        _bn_params(p, "bn_" + name, cout, dtype)
    _conv_params(p, "soft", 128, classes, gen, dtype, k=1)
    _conv_params(p, "sig", 128, 1, gen, dtype, k=1)
    return p


def trainable_names(p):
    return [k for k in p if not (k.endswith("moving_mean") or k.endswith("moving_variance"))]


def count_trainable(p):
    return sum(p[k].numel() for k in trainable_names(p))


# --------------------------------------------------------------------------------------------------
# VAE
# --------------------------------------------------------------------------------------------------
def tile_cond(cond, shape_dhw, reps=4):
    """Reshape((1,1,1,ncond)) + Lambda(K.tile, n=input_shape) (lattice_vae.py:167-168): the one-hot is
    broadcast to every voxel and repeated `reps`=input_shape[-1] times along the channel axis (R1)."""
    B = cond.shape[0]
    D, H, W = shape_dhw
    c = cond.reshape(B, 1, 1, 1, -1).expand(B, D, H, W, cond.shape[1])
    return c.repeat(1, 1, 1, 1, reps)


def vae_encoder(p, M, cond, eps, training, taps=None, bn_stats=None):
    """lattice_vae.py:160-195.  `eps` is the N(0,1) draw of `sampling` (:53-66), supplied explicitly."""
    taps = {} if taps is None else taps
    x = torch.cat([M, tile_cond(cond, M.shape[1:4], reps=M.shape[-1])], dim=-1)
    for i in range(1, 5):
        x = K.conv3d_same(x, p[f"enc_conv{i}/kernel"], p[f"enc_conv{i}/bias"])
        taps[f"enc_conv{i}"] = x
        x, mean, var = K.batchnorm(x, p[f"enc_bn{i}/gamma"], p[f"enc_bn{i}/beta"], p[f"enc_bn{i}/moving_mean"],
                                   p[f"enc_bn{i}/moving_variance"], training)
        if bn_stats is not None and training:
            bn_stats[f"enc_bn{i}"] = (mean.detach(), var.detach(), x.numel() // x.shape[-1])
        x = K.leaky_relu(x)
        taps[f"enc_bn{i}"] = x
        x = K.maxpool2(x)
        taps[f"enc_pool{i}"] = x
    x = K.leaky_relu(K.conv3d_same(x, p["enc_conv5/kernel"], p["enc_conv5/bias"]))
    taps["enc_conv5"] = x
    h = K.relu(K.dense(x.reshape(x.shape[0], -1), p["enc_dense/kernel"], p["enc_dense/bias"]))
    taps["enc_dense"] = h
    z_mean = K.dense(h, p["z_mean/kernel"], p["z_mean/bias"])
    z_log_var = K.dense(h, p["z_log_var/kernel"], p["z_log_var/bias"])
    z = z_mean + torch.exp(0.5 * z_log_var) * eps
    taps.update(z_mean=z_mean, z_log_var=z_log_var, z=z)
    return z_mean, z_log_var, z


def vae_decoder(p, z, cond, training, taps=None, bn_stats=None):
    """lattice_vae.py:197-230."""
    taps = {} if taps is None else taps
    h = K.dense(torch.cat([z, cond], dim=-1), p["dec_dense/kernel"], p["dec_dense/bias"])
    taps["dec_dense"] = h
    s = round((h.shape[1] // 4) ** (1 / 3))
    x = h.reshape(h.shape[0], s, s, s, 4)
    for i in range(1, 5):
        x = K.conv3d_same(x, p[f"dec_conv{i}/kernel"], p[f"dec_conv{i}/bias"])
        taps[f"dec_conv{i}"] = x
        x, mean, var = K.batchnorm(x, p[f"dec_bn{i}/gamma"], p[f"dec_bn{i}/beta"], p[f"dec_bn{i}/moving_mean"],
                                   p[f"dec_bn{i}/moving_variance"], training)
        if bn_stats is not None and training:
            bn_stats[f"dec_bn{i}"] = (mean.detach(), var.detach(), x.numel() // x.shape[-1])
        x = K.leaky_relu(x)
        taps[f"dec_bn{i}"] = x
        if i < 4:
            x = K.upsample2(x)
    x = K.conv3d_same(x, p["decoder_output/kernel"], p["decoder_output/bias"])
    taps["decoder_output"] = x
    x, mean, var = K.batchnorm(x, p["dec_bn5/gamma"], p["dec_bn5/beta"], p["dec_bn5/moving_mean"],
                               p["dec_bn5/moving_variance"], training)
    if bn_stats is not None and training:
        bn_stats["dec_bn5"] = (mean.detach(), var.detach(), x.numel() // x.shape[-1])
    x = K.relu(x)
    taps["x_hat"] = x
    return x


# --------------------------------------------------------------------------------------------------
# U-Net
# --------------------------------------------------------------------------------------------------
def _unet_block(p, name, x, training, taps, prefix):
    """Conv3D -> ReLU -> BatchNormalization, in that order (unet.py:276-278)."""
    a = K.relu(K.conv3d_same(x, p[name + "/kernel"], p[name + "/bias"]))
    taps[prefix + name] = a
    y, _, _ = K.batchnorm(a, p[f"bn_{name}/gamma"], p[f"bn_{name}/beta"], p[f"bn_{name}/moving_mean"],
                          p[f"bn_{name}/moving_variance"], training)
    taps[prefix + "bn_" + name] = y
    return a, y


def unet_prefix(p, x, training, taps=None, prefix="pm/"):
    """The part of the U-Net the perceptual loss evaluates: c1..c10 with taps after the ReLUs of
    c2,c4,c6,c10 (lattice_vae.py:257-263; R12).  BN follows the learning phase (R2)."""
    taps = {} if taps is None else taps
    feats = []
    _, y = _unet_block(p, "c1", x, training, taps, prefix)
    a, y = _unet_block(p, "c2", y, training, taps, prefix)
    feats.append(a)
    y = K.maxpool2(y)
    _, y = _unet_block(p, "c3", y, training, taps, prefix)
    a, y = _unet_block(p, "c4", y, training, taps, prefix)
    feats.append(a)
    y = K.maxpool2(y)
    _, y = _unet_block(p, "c5", y, training, taps, prefix)
    a, y = _unet_block(p, "c6", y, training, taps, prefix)
    feats.append(a)
    y = K.maxpool2(y)
    _, y = _unet_block(p, "c9", y, training, taps, prefix)
    a = K.relu(K.conv3d_same(y, p["c10/kernel"], p["c10/bias"]))
    taps[prefix + "c10"] = a
    feats.append(a)
    return feats


def unet_forward(p, x, training, taps=None):
    """unet.py:272-355 — returns (soft_logits, sig_logit); softmax/sigmoid are applied by the losses."""
    taps = {} if taps is None else taps
    pre = ""
    _, c1 = _unet_block(p, "c1", x, training, taps, pre)
    _, c2 = _unet_block(p, "c2", c1, training, taps, pre)
    _, c3 = _unet_block(p, "c3", K.maxpool2(c2), training, taps, pre)
    _, c4 = _unet_block(p, "c4", c3, training, taps, pre)
    _, c5 = _unet_block(p, "c5", K.maxpool2(c4), training, taps, pre)
    _, c6 = _unet_block(p, "c6", c5, training, taps, pre)
    _, c9 = _unet_block(p, "c9", K.maxpool2(c6), training, taps, pre)
    _, c10 = _unet_block(p, "c10", c9, training, taps, pre)
    _, c13 = _unet_block(p, "c13", torch.cat([c6, K.upsample2(c10)], dim=-1), training, taps, pre)
    _, c14 = _unet_block(p, "c14", c13, training, taps, pre)
    _, c15 = _unet_block(p, "c15", torch.cat([c4, K.upsample2(c14)], dim=-1), training, taps, pre)
    _, c16 = _unet_block(p, "c16", c15, training, taps, pre)
    _, c17 = _unet_block(p, "c17", torch.cat([c2, K.upsample2(c16)], dim=-1), training, taps, pre)
    _, c18 = _unet_block(p, "c18", c17, training, taps, pre)
    soft = K.conv3d_same(c18, p["soft/kernel"], p["soft/bias"])
    sig = K.conv3d_same(c18, p["sig/kernel"], p["sig/bias"])
    taps["soft_logits"] = soft
    taps["sig_logit"] = sig
    return soft, sig


# --------------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------------
def vae_dfc_loss(x, x_hat, z_mean, z_log_var, feats_x, feats_xhat, alpha=ALPHA, beta=BETA,
                 layer_weights=(1.0, 1.0, 1.0, 1.0)):
    """lattice_vae.py:232-255.  Returns (loss, pm, mse, kld) = the 4-tuple `train_on_batch` reports
    (lattice_vae.py:124-125): batch means of the per-sample terms."""
    mse = ((x - x_hat) ** 2).mean()
    kl = -0.5 * (1 + z_log_var - z_mean ** 2 - torch.exp(z_log_var)).sum(dim=-1)
    pm = 0.0
    for h1, h2, w in zip(feats_x, feats_xhat, layer_weights):
        pm = pm + w * ((h1.reshape(h1.shape[0], -1) - h2.reshape(h2.shape[0], -1)) ** 2).mean(dim=-1)
    loss = (mse + alpha * pm + beta * kl).mean()
    return loss, pm.mean(), mse, kl.mean()


def vae_dfc_step(p_vae, p_pm, M, cond, eps, training=True, pm_bn_training=None, alpha=ALPHA, beta=BETA, taps=None,
                 bn_stats=None):
    """One forward evaluation of the compiled model of lattice_vae.py:127-145 (target = input, :296-298)."""
    taps = {} if taps is None else taps
    pm_bn_training = training if pm_bn_training is None else pm_bn_training
    z_mean, z_log_var, z = vae_encoder(p_vae, M, cond, eps, training, taps, bn_stats)
    x_hat = vae_decoder(p_vae, z, cond, training, taps, bn_stats)
    feats_x = unet_prefix(p_pm, M, pm_bn_training, taps, prefix="pm_x/")
    feats_xh = unet_prefix(p_pm, x_hat, pm_bn_training, taps, prefix="pm_xhat/")
    return vae_dfc_loss(M, x_hat, z_mean, z_log_var, feats_x, feats_xh, alpha, beta), x_hat


def weighted_cce(soft_logits, species, weight=95.0, eps=1e-7):
    """unet.py:196-221 on a softmax output: renormalise, clip to [1e-7, 1-1e-7] (zero gradient outside),
    -sum_c y_c log p_c w_c, mean over voxels then batch.  `weight` is the scalar num_classes the reference
    actually passes (unet.py:254) or a (C,) vector.  `species`: integer labels."""
    p = torch.softmax(soft_logits, dim=-1)
    p = p / p.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, eps, 1 - eps)
    logp = torch.log(p)
    picked = torch.gather(logp, -1, species.long().unsqueeze(-1)).squeeze(-1)
    if torch.is_tensor(weight) and weight.dim() == 1:
        wv = weight[species.long()]
    else:
        wv = weight
    per_voxel = -(picked * wv)
    return per_voxel.reshape(per_voxel.shape[0], -1).mean(dim=1).mean()


def sigmoid_bce(sig_logit, mask):
    """Keras binary_crossentropy on the sigmoid head, logits form (R10): mean over everything."""
    x = sig_logit.squeeze(-1)
    t = mask.to(x.dtype)
    return (torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))).mean()


def unet_metrics(soft_logits, species, classes=95, eps=1e-7):
    """f1_m and wr_m of unet.py:159-193 on the softmax head (batch-level sums of rounded products)."""
    p = torch.softmax(soft_logits, dim=-1)
    y = torch.nn.functional.one_hot(species.long(), classes).to(p.dtype)
    rnd = lambda t: torch.round(torch.clamp(t, 0, 1))
    tp = rnd(y * p).sum()
    possible = rnd(y).sum()
    predicted = rnd(p).sum()
    recall = tp / (possible + eps)
    precision = tp / (predicted + eps)
    f1 = 2 * (precision * recall) / (precision + recall + eps)
    w = torch.ones(classes, dtype=p.dtype)
    w[0] = 0
    wr = rnd(w * y * p).sum() / (rnd(w * y).sum() + eps)
    return f1, wr


def unet_loss(p, x, species, training=True, weight=95.0, taps=None):
    """Total loss of the compiled U-Net (unet.py:252-259): soft + sig with unit loss weights.
    Returns [loss, soft_loss, sig_loss, f1_m, wr_m] — the order Keras reports."""
    soft, sig = unet_forward(p, x, training, taps)
    ls = weighted_cce(soft, species, weight)
    lb = sigmoid_bce(sig, species != 0)
    f1, wr = unet_metrics(soft.detach(), species)
    return [ls + lb, ls, lb, f1, wr], soft, sig


# --------------------------------------------------------------------------------------------------
# optimiser
# --------------------------------------------------------------------------------------------------
class KerasAdam:
    """keras.optimizers.Adam (2.3.1): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps), eps=1e-7 (R11)."""

    def __init__(self, lr, beta_1=0.9, beta_2=0.999, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for k, g in grads.items():
            if k not in self.m:
                self.m[k] = torch.zeros_like(params[k])
                self.v[k] = torch.zeros_like(params[k])
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            params[k] = params[k] - lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)
        return params


def vae_train_step(p_vae, p_pm, opt, M, cond, eps):
    """train_on_batch([M,cond], M) (lattice_vae.py:296-298): forward, backward w.r.t. the VAE's trainable
    weights only (pm is frozen — SURVEY A5), Adam update, BN moving-average update.  Returns the metrics."""
    names = trainable_names(p_vae)
    leaves = {k: p_vae[k].detach().clone().requires_grad_(True) for k in names}
    pv = dict(p_vae)
    pv.update(leaves)
    bn_stats = {}
    (loss, pm, mse, kl), x_hat = vae_dfc_step(pv, p_pm, M, cond, eps, training=True, bn_stats=bn_stats)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    grads = dict(zip(names, grads))
    new = opt.step({k: p_vae[k] for k in names}, grads)
    for k in names:
        p_vae[k] = new[k].detach()
    for bn, (mean, var, n) in bn_stats.items():
        mm, mv = K.bn_moving_update(p_vae[bn + "/moving_mean"], p_vae[bn + "/moving_variance"], mean, var, n)
        p_vae[bn + "/moving_mean"], p_vae[bn + "/moving_variance"] = mm, mv
    return [float(loss), float(pm), float(mse), float(kl)], grads
