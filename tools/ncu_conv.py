"""ncu target: the dominant conv shapes of the default model (one launch each after warm-up).
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 6 -c 3 -o gpurun_out/prof_conv python tools/ncu_conv.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops
B, H, W = 32, 448, 576
shapes = [(64, 64, 3, True), (32, 32, 3, True), (64, 32, 1, True)]
xs = [(torch.randn(B, H, W, ci, device="cuda").to(torch.bfloat16), torch.randn(co, ci, k, k, device="cuda") * 0.05, st)
      for ci, co, k, st in shapes]
for rep in range(3):      # launches 0..5 warm-up, 6..8 profiled (64->64 3x3 stats, 32->32 3x3 stats, 64->32 1x1 stats)
    for x, w, st in xs:
        ops.conv_tc(x, w, None, stats=st)
torch.cuda.synchronize()
print("ok")
