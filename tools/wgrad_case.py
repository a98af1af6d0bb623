"""Run ONE filter-gradient shape a few times (for ncu / timing).  usage: wgrad_case.py B D cin cout [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icsg3d_b200 import ops
B, D, cin, cout = map(int, sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
x = torch.randn(B, D, D, D, cin, device="cuda").to(torch.bfloat16)
dy = torch.randn(B, D, D, D, cout, device="cuda").to(torch.bfloat16)
ws = torch.empty(max(ops.conv3d_k3_wgrad_workspace_bytes(B, D, cin, cout), 16), dtype=torch.uint8, device="cuda")
out = torch.empty(27, cin, cout, dtype=torch.float32, device="cuda")
for _ in range(2):
    ops.conv3d_k3_wgrad(x, dy, out=out, ws=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.conv3d_k3_wgrad(x, dy, out=out, ws=ws)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"wgrad B={B} D={D} {cin}->{cout}: {ms*1e3:.1f} us  {2.0*B*D**3*27*cin*cout/ms/1e9:.1f} TF/s")
