"""Achieved HBM bandwidth of the BatchNorm kernels (forward apply, backward reduce, backward apply)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, depth_b200
from depth_b200 import ops, _lib as L
lib = L.lib()
BF = torch.bfloat16


def t(fn, reps=5):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for (B, H, W, C) in [(32, 448, 576, 64), (32, 448, 576, 32), (32, 448, 576, 16), (32, 224, 288, 144), (32, 112, 144, 192), (32, 28, 36, 816)]:
    npix = B * H * W
    x = torch.randn(B, H, W, C, device="cuda").to(BF)
    gy = torch.randn(B, H, W, C, device="cuda").to(BF)
    y = torch.empty_like(x); dx = torch.empty_like(x)
    ss = torch.rand(2, C, device="cuda"); save = torch.rand(2, C, device="cuda") + 0.5
    gamma = torch.rand(C, device="cuda")
    nb = lib.dp_chan_reduce_blocks()
    part = torch.empty(nb, 2, C, device="cuda"); red = torch.empty(2, C, device="cuda")
    dg = torch.empty(C, device="cuda"); db = torch.empty(C, device="cuda")
    st = L.stream()
    gb = npix * C * 2 / 1e9
    ms = t(lambda: L.check(lib.dp_bn_apply(L.ptr(x), C, L.ptr(ss), None, 0, None, None, 0, npix, C, 1, L.ptr(y), C, st)))
    print(f"{H}x{W}x{C}: bn_apply            {ms:6.3f} ms  {2 * gb / ms * 1e3:7.0f} GB/s")
    ms = t(lambda: L.check(lib.dp_chan_reduce(1, L.ptr(x), C, None, 0, None, 0, None, npix, C, L.ptr(part), st)))
    print(f"{H}x{W}x{C}: chan_reduce<1>      {ms:6.3f} ms  {1 * gb / ms * 1e3:7.0f} GB/s")
    ms = t(lambda: L.check(lib.dp_chan_reduce(2, L.ptr(x), C, L.ptr(gy), C, None, 0, L.ptr(ss), npix, C, L.ptr(part), st)))
    print(f"{H}x{W}x{C}: chan_reduce<2> rec  {ms:6.3f} ms  {2 * gb / ms * 1e3:7.0f} GB/s")
    ms = t(lambda: L.check(lib.dp_bn_bwd_apply(L.ptr(gy), C, None, 0, L.ptr(ss), L.ptr(x), C, L.ptr(red), L.ptr(save), L.ptr(gamma),
                                              float(npix), 1, npix, C, L.ptr(dx), C, None, 0, L.ptr(dg), L.ptr(db), 0, st)))
    print(f"{H}x{W}x{C}: bn_bwd_apply rec    {ms:6.3f} ms  {3 * gb / ms * 1e3:7.0f} GB/s")
