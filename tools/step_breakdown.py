"""Per-kernel GPU-time breakdown of one eager train step (torch.profiler / CUPTI), grouped by kernel name.
Usage (GPU box): python tools/step_breakdown.py [--batch 32] > gpurun_out/breakdown.txt"""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import depth_b200
from depth_b200 import config as fx

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
a = ap.parse_args()
if os.environ.get("DP_NO_BN_FUSION"):
    from depth_b200 import ops as _ops
    _ops.Fusion.prologue = _ops.Fusion.backward = False
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
cfg = fx.loss_config()
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, fused=True)
x, t = bench.synthetic_batch(a.batch, 1234)
x, t = x.to(dev), t.to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    out = model(x).unsqueeze(1)
    loss, parts = depth_b200.combined_loss(out, t, cfg, rgb=x)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity, record_function
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    with record_function("fwd"):
        out = model(x).unsqueeze(1)
        loss, parts = depth_b200.combined_loss(out, t, cfg, rgb=x)
    torch.cuda.synchronize()
    with record_function("bwd"):
        loss.backward()
    torch.cuda.synchronize()
    with record_function("opt"):
        opt.step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0:
        name = ev.name
        agg[name][0] += ev.device_time
        agg[name][1] += 1
        tot += ev.device_time
print(f"total GPU kernel time {tot / 1e3:.2f} ms  (batch {a.batch})")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
    print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  n={n:4d}  {name[:150]}")
