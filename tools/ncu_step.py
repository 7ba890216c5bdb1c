"""ncu target: one train step (B = 32) of the default model, issued eagerly with the same body GraphedTrainStep captures
(side-stream weight gradients included), after an `arange` marker kernel that tools/launch_summary.py keys on.  The
graph replay launches exactly these kernels; ncu itself fails with LaunchFailed on the first conv_tc_kernel launched
under stream capture in this build (cause not found - the same launches profile fine outside capture), so the list is
taken from the eager issue of the step.
  python tools/ncu_step.py > gpurun_out/step_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py
then: python tools/launch_summary.py gpurun_out/launches.csv profiles/ncu_launches_<tag>_summary.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import depth_b200
from depth_b200 import config as fx, ops, util

B = int(os.environ.get("B", "32"))
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, fused=True, capturable=True)
x, t = bench.synthetic_batch(B, 1234)
x, t = x.to(dev), t.to(dev)
cfg = fx.loss_config()


def body():
    for p in model.parameters():
        p.grad = None
    pred = model(x).unsqueeze(1)
    total, out = util.combined_loss_device(pred, t, cfg, rgb=x)
    ops.side_enable(True)
    try:
        total.backward()
        ops.side_join()
    finally:
        ops.side_enable(False)
    opt.step()
    return out


for _ in range(2):
    body()
torch.cuda.synchronize()
marker = torch.arange(7, device=dev)          # a kernel name the marker owns: wait - the ViT stand-in uses arange too,
marker = torch.tril(torch.ones(3, 3, device=dev))   # so the summary keys on `triu_tril`, which appears nowhere else
torch.cuda.synchronize()
out = body()
torch.cuda.synchronize()
print("loss", float(out[depth_b200._lib.L_SI]), "build", depth_b200._lib.build_id())
