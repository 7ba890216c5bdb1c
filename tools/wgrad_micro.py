"""Times the weight-gradient shapes of the default model against their HBM / tensor floors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops
HBM, TF = 6549.8e9, 1383.7e12
shapes = [(32, 448, 576, 64, 64, 3), (32, 448, 576, 32, 32, 3), (32, 448, 576, 16, 16, 3), (32, 448, 576, 64, 32, 3),
          (32, 448, 576, 32, 16, 3), (32, 448, 576, 64, 32, 1), (32, 448, 576, 32, 16, 1),
          (32, 112, 144, 64, 64, 3), (32, 64, 80, 128, 128, 3), (32, 56, 72, 128, 128, 3), (32, 32, 40, 128, 128, 3),
          (32, 28, 36, 256, 256, 3), (32, 16, 20, 128, 128, 3), (32, 14, 18, 512, 512, 3), (32, 14, 18, 384, 512, 3),
          (32, 16, 20, 384, 512, 1), (32, 8, 10, 128, 128, 3)]
for B, H, W, cin, cout, ks in shapes:
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    g = torch.randn(B, H, W, cout, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        ops._wgrad_tc(x, g, cin, cout, ks)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        ops._wgrad_tc(x, g, cin, cout, ks)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    fl = 2.0 * B * H * W * cin * cout * ks * ks
    by = 2.0 * B * H * W * (cin + cout)
    floor = max(fl / TF, by / HBM) * 1e3
    print(f"wgrad {H}x{W} {cin:3d}->{cout:3d} k{ks}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TF/s  {by / ms / 1e6:7.1f} GB/s  floor {floor:.3f} ms "
          f"({100 * floor / ms:.0f}% of {'tensor' if fl / TF > by / HBM else 'hbm'} roofline)")
