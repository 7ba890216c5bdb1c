import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops
B, H, W, cin, cout = [int(v) for v in os.environ.get("SHAPE", "32,14,18,384,1392").split(",")]
x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
w = torch.randn(cout, cin, 1, 1, device="cuda") * 0.05
y = ops.conv_tc(x, w, None)
torch.cuda.synchronize()
ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float()).permute(0, 2, 3, 1)
print("ok", float((y.float() - ref).abs().max()), float(ref.abs().max()))
