"""Run ONE conv shape a few times (for ncu / timing).  usage: conv_case.py B D cin cout [iters] ; ICSG3D_CONV_IMPL=v1|halo"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icsg3d_b200 import ops
B, D, cin, cout = map(int, sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
x = torch.randn(B, D, D, D, cin, device="cuda").to(torch.bfloat16)
w = (torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5)
wp = ops.pack_conv_w_fprop(w)
bias = torch.zeros(cout, device="cuda")
y = torch.empty(B, D, D, D, cout, dtype=torch.bfloat16, device="cuda")
# the engines always pass their split-K workspace (used by the per-tap kernel on layers with few tiles)
ws = torch.empty(max(ops.conv3d_k3_workspace_bytes(B, D, cin, cout), 16), dtype=torch.uint8, device="cuda")
_conv = ops.conv3d_k3
ops.conv3d_k3 = lambda *a, **k: _conv(*a, ws=ws, **k)
for _ in range(2):
    ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.conv3d_k3(x, wp, bias, act=ops.ACT_RELU, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"impl={os.environ.get('ICSG3D_CONV_IMPL','halo')} B={B} D={D} {cin}->{cout}: {ms*1e3:.1f} us  {2.0*B*D**3*27*cin*cout/ms/1e9:.1f} TF/s")
