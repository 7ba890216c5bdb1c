"""Pointwise (1x1) trunk convolutions with BatchNorm statistics: GPU time per launch (graph-timed) against the HBM floor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops
HBM = 6549.8e9
torch.cuda.set_stream(torch.cuda.Stream())
B = 32
SHAPES = [(28, 36, 136, 816), (28, 36, 816, 136), (14, 18, 232, 1392), (14, 18, 1392, 232), (112, 144, 32, 192),
          (112, 144, 192, 32), (56, 72, 48, 288), (56, 72, 288, 48), (28, 36, 96, 576), (28, 36, 576, 96), (224, 288, 24, 144)]
for (H, W, ci, co) in SHAPES:
    x = torch.randn(B, H, W, ci, device="cuda").to(torch.bfloat16)
    wp = (torch.randn(1, co, ci, device="cuda") * 0.05).to(torch.bfloat16)
    res = []
    for stats in (True, False):
        fn = lambda: ops._conv_raw(x, wp, co, 1, stats=stats)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10): fn()
        g.replay(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        res.append(s.elapsed_time(e) / 10 * 1e3)
    by = 2.0 * B * H * W * (ci + co)
    print(f"1x1 {ci:5d}->{co:5d} @{H}x{W}: +stats {res[0]:7.1f} us  plain {res[1]:7.1f} us  (HBM floor {by / HBM * 1e6:5.1f} us)", flush=True)
