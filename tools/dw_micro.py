"""Depthwise conv kernels of the EfficientNet trunk against their HBM floors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import ops
HBM = 6549.8e9
torch.cuda.set_stream(torch.cuda.Stream())     # everything off the legacy stream: the timing loop is graph-captured
SHAPES = [(32, 56, 72, 288, 5, 1), (32, 28, 36, 816, 5, 1), (32, 14, 18, 1392, 5, 1), (32, 112, 144, 192, 3, 1),
          (32, 224, 288, 144, 3, 2), (32, 14, 18, 1392, 3, 1)]
if len(sys.argv) > 1:
    SHAPES = [SHAPES[int(sys.argv[1])]]
EAGER = len(sys.argv) > 2      # ncu: plain eager launches
for (B, H, W, C, K, S) in SHAPES:
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(C, 1, K, K, device="cuda") * 0.2).requires_grad_(True)
    p = K // 2
    Ho, Wo = (H + 2 * p - K) // S + 1, (W + 2 * p - K) // S + 1
    def fwd():
        return ops.dwconv(x, w, S, p, p, Ho, Wo, stats=True)[0]
    y = fwd(); g = torch.randn_like(y)
    if EAGER:
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        continue

    def t(fn, reps=10):
        """GPU time per call: the calls are captured in a CUDA graph (the step runs as a graph; eager timing of these
        small kernels measures Python dispatch instead)"""
        for _ in range(3): fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(reps): fn()
        gr.replay(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        gr.replay()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / reps * 1e3
    tf = t(fwd)
    xd, wd = x.detach(), w.detach()
    def fbx():
        yy = ops.dwconv(x, wd, S, p, p, Ho, Wo, stats=True)[0]; torch.autograd.grad(yy, (x,), g)
    def fbw():
        yy = ops.dwconv(xd, w, S, p, p, Ho, Wo, stats=True)[0]; torch.autograd.grad(yy, (w,), g)
    tdx = t(fbx) - tf
    tdw = t(fbw) - tf
    by = 2.0 * B * (H * W + Ho * Wo) * C
    print(f"dw k{K} s{S} {H}x{W}x{C}: fwd {tf:7.1f} us (floor {by / HBM * 1e6:5.1f})  dgrad {tdx:7.1f} us  wgrad {tdw:7.1f} us "
          f"(floor {by / HBM * 1e6:5.1f} each)", flush=True)
