"""Times the BatchNorm-fused conv launches against their unfused counterparts at the benched full-resolution shapes."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops, _lib as L
B, H, W = 32, 448, 576
HBM = 6549.8e9


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def pack(w):
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous().to(torch.bfloat16)


shapes = [(64, 64, 3), (32, 32, 3), (16, 16, 3), (64, 32, 3), (32, 16, 3), (16, 32, 3), (32, 64, 3), (64, 32, 1), (32, 16, 1)]
if len(sys.argv) > 1:
    B = int(sys.argv[1])
if len(sys.argv) > 2:          # single shape "cin,cout,ks" (ncu captures)
    shapes = [tuple(int(v) for v in sys.argv[2].split(","))]
for cin, cout, ks in shapes:
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    w = pack(torch.randn(cout, cin, ks, ks, device="cuda") * 0.05)
    ss = torch.stack([torch.rand(cin, device="cuda") + 0.5, torch.randn(cin, device="cuda")]).contiguous()
    ssm = torch.stack([torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")]).contiguous()
    c = torch.randn(B, H, W, cout, device="cuda").to(torch.bfloat16)
    g = torch.randn(B, H, W, cout, device="cuda").to(torch.bfloat16)
    t_plain = timeit(lambda: ops._conv_raw(x, w, cout, ks, stats=True))
    t_pre = timeit(lambda: ops._conv_raw(x, w, cout, ks, stats=True, pre=(ss, 1)))
    caps = L.lib().dp_conv2d_tc_caps(B, H, W, cin, cout, ks)
    t_mask = timeit(lambda: ops._conv_raw(x, w, cout, ks, mask=(c, ssm, 1))) if caps & 2 else float("nan")
    t_res = timeit(lambda: ops._conv_raw(x, w, cout, ks, res=c))
    t_bn = timeit(lambda: ops._bn_apply_raw(x, ss, 1))
    t_red = timeit(lambda: ops._bn_reduce(c, g, None, ssm, 1))
    t_wg = timeit(lambda: ops._wgrad_raw(x, g, cin, cout, ks))
    t_wgp = timeit(lambda: ops._wgrad_raw(x, g, cin, cout, ks, pre=(ss, 1)))
    by = 2.0 * B * H * W * (cin + cout)
    print(f"{cin:3d}->{cout:3d} k{ks}: conv {t_plain:.3f} (floor {by / HBM * 1e3:.3f})  +res {t_res:.3f}  +pre {t_pre:.3f}  +mask {t_mask:.3f} | "
          f"bn_apply(in) {t_bn:.3f} reduce(out) {t_red:.3f} | wgrad {t_wg:.3f} +pre {t_wgp:.3f}")
