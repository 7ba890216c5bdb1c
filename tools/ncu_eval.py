"""ncu target: the streaming evaluation-metrics kernel (exact and fast-math) at the bench's eval shape (650 x 448 x 576):
ncu --set full --clock-control none --import-source on -k regex:eval_stream_kernel -s 4 -c 2 python tools/ncu_eval.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
EB, H, W = 650, 448, 576
g = torch.Generator(device="cuda").manual_seed(7)
tt = torch.rand(EB, 1, H, W, device="cuda", generator=g) * 9.9 + 0.1
pp = tt * torch.exp(0.1 * torch.randn(EB, 1, H, W, device="cuda", generator=g)) * 1.3
for _ in range(3):
    depth_b200.evaluation_metrics(pp, tt)
    depth_b200.evaluation_metrics(pp, tt, fast_math=True)
torch.cuda.synchronize()
print("ok")
