"""torch.profiler kernel-level breakdown of one train step (and of the eval reductions)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import depth_b200
from depth_b200 import distributed as D
from oracle import fixtures as fx
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
cfg = fx.loss_config()
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, fused=True)
red = D.GradientAllReducer(model.parameters(), 1)
x, t = bench.synthetic_batch(B, 1234)
x, t = x.to(dev), t.to(dev)
def step():
    red.zero()
    out = model(x).unsqueeze(1)
    loss, parts = depth_b200.combined_loss(out, t, cfg, rgb=x)
    loss.backward(); red.reduce(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
import time
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print("wall ms/step", (time.perf_counter() - t0) * 1e3)
tt = torch.rand(128, 1, 448, 576, device=dev) * 9.9 + 0.1
pp = tt * 1.3
for _ in range(3): depth_b200.evaluation_metrics(pp, tt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    depth_b200.evaluation_metrics(pp, tt); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
