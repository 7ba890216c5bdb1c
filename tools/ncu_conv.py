"""a handful of full-resolution conv / wgrad launches for `ncu --set full` (one GPU, short)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_b200
from depth_b200 import ops
B, H, W = 8, 448, 576
torch.manual_seed(0)
for cin, cout in [(64, 64), (32, 32)]:
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    g = torch.randn(B, H, W, cout, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    for _ in range(2):
        y = ops.conv_tc(x, w, None, stats=True)
        dw = ops._wgrad_tc(x, g, cin, cout, 3)
    torch.cuda.synchronize()
print("ok")
