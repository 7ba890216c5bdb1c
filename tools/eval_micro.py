"""Times the evaluation-reduction kernels one by one (CUDA events) at the bench's eval shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, depth_b200
from depth_b200 import util, _lib as L
EB, H, W = int(os.environ.get("EB", "650")), 448, 576
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
tt = torch.rand(EB, 1, H, W, device=dev, generator=g) * 9.9 + 0.1
pp = tt * torch.exp(0.1 * torch.randn(EB, 1, H, W, device=dev, generator=g)) * 1.3
pp[:, :, 100:140, 200:300] = 0.0


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


ps = util._Pass(pp, tt, None, L.F_SI | L.F_ABSREL, 1e-6)
print("moments pass      ms", t(lambda: util._Pass(pp, tt, None, L.F_SI | L.F_ABSREL, 1e-6)))
print("delta counts pass ms", t(lambda: ps.counts([1.05, 1.05 ** 2, 1.05 ** 3], aligned=True)))
px = EB * H * W
for name, fm in (("default (lean)", None), ("exact IEEE", False), ("MUFU", True)):
    ms = t(lambda: depth_b200.evaluation_metrics(pp, tt, fast_math=fm))
    print(f"evaluation_metrics {name:15s} {ms:.4f} ms  {px / ms / 1e6:7.1f} Gpx/s  {8 * px / ms / 1e6:7.1f} GB/s  out {[round(v, 7) for v in depth_b200.evaluation_metrics(pp, tt, fast_math=fm).tolist()]}")
scratch = torch.empty_like(tt)
print("copy (read+write 8B/px) ms", t(lambda: scratch.copy_(pp)))
