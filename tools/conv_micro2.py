"""EfficientNet trunk 1x1 shapes: time with / without the BN-statistics epilogue."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops
HBM, TF = 6549.8e9, 1383.7e12
shapes = [(32, 112, 144, 32, 192), (32, 224, 288, 24, 144), (32, 28, 36, 136, 816), (32, 28, 36, 96, 576), (32, 56, 72, 48, 288),
          (32, 14, 18, 232, 1392), (32, 14, 18, 1392, 232), (32, 28, 36, 816, 136), (32, 112, 144, 192, 32), (32, 224, 288, 144, 32)]
for B, H, W, cin, cout in shapes:
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, 1, 1, device="cuda") * 0.05
    res = []
    for st in (False, True):
        for _ in range(3):
            ops.conv_tc(x, w, None, stats=st)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(2e6))
        s.record()
        for _ in range(10):
            ops.conv_tc(x, w, None, stats=st)
        e.record(); torch.cuda.synchronize()
        res.append(s.elapsed_time(e) / 10)
    fl = 2.0 * B * H * W * cin * cout
    by = 2.0 * B * H * W * (cin + cout)
    floor = max(fl / TF, by / HBM) * 1e3
    print(f"1x1 {H}x{W} {cin:4d}->{cout:4d}: plain {res[0]*1e3:7.1f} us  +stats {res[1]*1e3:7.1f} us  floor {floor*1e3:6.1f} us  ({by/res[1]/1e6:6.0f} GB/s with stats)")
