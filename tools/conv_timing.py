"""Per-role cycle counters of conv_tc_kernel (needs a build with -DDP_CONV_TIMING: DP_EXTRA_FLAGS=-DDP_CONV_TIMING
python monocular-depth-estimation-cil_b200/build.py -f)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops, _lib as L
B, H, W = 32, 448, 576
dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
L.lib().dp_debug_set_buffer(L.ptr(dbg))
for cin, cout, ks, stats in [(64, 64, 3, False), (64, 64, 3, True), (32, 32, 3, False), (16, 16, 3, False), (64, 32, 1, False), (32, 16, 1, False)]:
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, ks, ks, device="cuda") * 0.05
    for _ in range(2):
        dbg.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.conv_tc(x, w, None, stats=stats); e.record(); torch.cuda.synchronize()
    d = dbg.tolist()
    n = max(d[7], 1)
    print(f"cin {cin} cout {cout} ks {ks} stats {stats}: {s.elapsed_time(e):.3f} ms; tiles/CTA {d[7]}; total/tile {d[3]/n:.0f} cyc | issuer0 (its {d[4]}): wait tmem-empty {d[0]/n:.0f} wait-full {d[1]/n:.0f} issue {d[2]/n:.0f} | epilogue: wait {d[5]/n:.0f} work {d[6]/n:.0f}")
L.lib().dp_debug_set_buffer(None)
