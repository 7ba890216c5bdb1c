import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops, _lib as L
B, H, W = 32, 448, 576
dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
L.lib().dp_debug_set_buffer(L.ptr(dbg))
for cin, cout, ks in [(32, 32, 3), (64, 64, 3), (32, 16, 1), (16, 16, 3)]:
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    g = torch.randn(B, H, W, cout, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        dbg.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops._wgrad_tc(x, g, cin, cout, ks); e.record(); torch.cuda.synchronize()
    d = dbg.tolist()
    print(f"cin {cin} cout {cout} ks {ks}: {s.elapsed_time(e):.3f} ms; block0: tiles {d[4]} total {d[3]} cyc; producer wait-empty {d[0]}; "
          f"mma wait-full {d[1]}; mma issue {d[2]}; per tile total {d[3]/max(d[4],1):.0f} wait {d[1]/max(d[4],1):.0f} issue {d[2]/max(d[4],1):.0f}")
L.lib().dp_debug_set_buffer(None)
