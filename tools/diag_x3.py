import sys, torch
sys.path.insert(0, '.')
from icsg3d_b200 import ops
from icsg3d_b200.engine_x3 import UNetForwardX3
from icsg3d_b200.params import ParamStore, unet_specs
from oracle import nets, keras_ops as K
from tests.util import rel_l2, synthetic_batch
# single conv vs fp64
g = torch.Generator().manual_seed(0)
B, D, cin, cout = 2, 16, 32, 64
x = torch.randn(B, D, D, D, cin, generator=g); w = torch.randn(3, 3, 3, cin, cout, generator=g) / (27 * cin) ** 0.5; b = torch.randn(cout, generator=g)
ref64 = K.conv3d_same(x.double(), w.double(), b.double()); ref32 = K.conv3d_same(x, w, b)
for fmt in (0, 1):
    ops.SPLIT_FMT = fmt
    x3 = torch.zeros(B, D, D, D, 3 * cin, dtype=torch.bfloat16, device="cuda"); ops.f32_to_split3(x.cuda(), cin, x3, cin)
    wp = ops.pack_conv_w_fprop_x3(w.cuda()); y = ops.conv3d_k3(x3, wp, b.cuda(), out_dtype=torch.float32, split=True)
    print("fmt", fmt, "conv vs fp64", rel_l2(y, ref64), "cpu fp32 vs fp64", rel_l2(ref32, ref64))
ops.SPLIT_FMT = 1
pp = ParamStore(unet_specs(4, 95), "cuda", with_grads=False, with_adam=False).init(5)
g = torch.Generator().manual_seed(11)
for k, v in pp.p.items():
    if k.endswith("moving_mean"): v.copy_(torch.rand(v.shape, generator=g) * 0.2)
    elif k.endswith("moving_variance"): v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.05)
    elif k.endswith("beta"): v.copy_(torch.randn(v.shape, generator=g) * 0.1)
M, _, _ = synthetic_batch(1, d=32, seed=2)
pu = {k: torch.from_numpy(v) for k, v in pp.to_dict().items()}
taps = {}
with torch.no_grad():
    soft, sig = nets.unet_forward(pu, M, training=False, taps=taps)
un = UNetForwardX3(1, d=32, params=pp)
logits, argmax, sigp = un.predict(M)
for k, v in un.taps.items():
    print(k, f"{rel_l2(v, taps[k]):.2e}", "max|a|", float(v.abs().max()))
print("logits", rel_l2(logits[..., :95], soft))
