"""Per-step end-to-end times with the copy-stream H2D overlap on, to see whether its slow mode is stable within a process."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import depth_b200
from depth_b200 import config as fx
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, fused=True, capturable=True)
xh, th = bench.synthetic_batch(32, 1234)
xh, th = xh.pin_memory(), th.pin_memory()
g = depth_b200.GraphedTrainStep(model, opt, fx.loss_config(), xh.to(dev), th.to(dev), use_rgb=True, world=1, warmup=3,
                                overlap_h2d=int(os.environ.get("OVERLAP", "1")) == 1)
for _ in range(3):
    g()
torch.cuda.synchronize()
if os.environ.get("SAMPLER"):
    sm = bench.ClockSampler(0)
    sm.start()
    for _ in range(16):
        g()
    torch.cuda.synchronize()
    print("clocks", sm.stop())
ts = []
g(xh, th); g.loss_dict(lag=0)
for i in range(40):
    t0 = time.perf_counter()
    g(xh, th)
    g.loss_dict(lag=1)
    ts.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
print("ms per step:", " ".join(f"{t:.1f}" for t in ts))
print("mean of last 30: %.2f" % (sum(ts[10:]) / 30))
