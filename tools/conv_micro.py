"""Times the full-resolution convolution shapes of the default model one by one against their HBM / tensor floors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops
B, H, W = 32, 448, 576
HBM, TF = 6549.8e9, 1383.7e12
shapes = [(64, 64, 3, False), (64, 64, 3, True), (32, 32, 3, False), (32, 32, 3, True), (16, 16, 3, True),
          (64, 32, 3, True), (64, 32, 1, True), (32, 16, 3, True), (32, 16, 1, True), (32, 64, 3, False), (16, 32, 3, False)]
small = [(32, 112, 144, 64, 64, 3), (32, 64, 80, 128, 128, 3), (32, 56, 72, 128, 128, 3), (32, 28, 36, 256, 256, 3),
         (32, 14, 18, 512, 512, 3)]


def run(Bq, Hq, Wq, cin, cout, ks, stats):
    x = torch.randn(Bq, Hq, Wq, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, ks, ks, device="cuda") * 0.05
    for _ in range(3):
        ops.conv_tc(x, w, None, stats=stats)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        ops.conv_tc(x, w, None, stats=stats)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    fl = 2.0 * Bq * Hq * Wq * cin * cout * ks * ks
    by = 2.0 * Bq * Hq * Wq * (cin + cout)
    floor = max(fl / TF, by / HBM) * 1e3
    print(f"{Hq}x{Wq} {cin:3d}->{cout:3d} k{ks} stats={int(stats)}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TF/s  {by / ms / 1e6:7.1f} GB/s  "
          f"floor {floor:.3f} ms ({100 * floor / ms:.0f}% of {'tensor' if fl / TF > by / HBM else 'hbm'} roofline)")


for cin, cout, ks, st in shapes:
    run(B, H, W, cin, cout, ks, st)
for Bq, Hq, Wq, cin, cout, ks in small:
    run(Bq, Hq, Wq, cin, cout, ks, False)
