"""Times the bilinear resize forward / backward kernels per shape against their HBM floor (read in + write out)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops

dev = torch.device("cuda", 0)
BF = torch.bfloat16


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


shapes = [(32, 224, 288, 64, 448, 576, True), (32, 224, 288, 32, 448, 576, False), (32, 112, 144, 128, 224, 288, True),
          (32, 128, 160, 64, 224, 280, True), (8, 224, 288, 256, 448, 576, True), (8, 448, 576, 128, 896, 1152, True)]
for B, Hi, Wi, C, Ho, Wo, al in shapes:
    x = torch.randn(B, Hi, Wi, C, device=dev).to(BF).requires_grad_(True)
    y = ops.resize(x, (Ho, Wo), al)
    g = torch.randn_like(y)
    byts = (x.numel() + y.numel()) * 2
    f = t(lambda: ops.resize(x.detach(), (Ho, Wo), al))
    def bwd():
        x.grad = None
        y.backward(g, retain_graph=True)
    b = t(bwd)
    c = t(lambda: g.clone())
    print(f"B{B} {Hi}x{Wi}x{C} -> {Ho}x{Wo} align={al}: fwd {f:.3f} ms {byts / f / 1e6:.0f} GB/s | bwd {b:.3f} ms {byts / b / 1e6:.0f} GB/s"
          f" | clone of out {c:.3f} ms {2 * y.numel() * 2 / c / 1e6:.0f} GB/s")
