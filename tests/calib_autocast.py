"""Calibration: how far does PLAIN PYTORCH bf16 autocast drift from the fp32 oracle on the same full-model
train step?  (Used to state the bf16 tolerance of tests/test_modules_gpu.py::test_full_model_*.)"""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_b200
from oracle import cases, fixtures as fx, losses as ol
from tests.test_modules_gpu import _full, rel_err, rel_l2
which = sys.argv[1] if len(sys.argv) > 1 else "semantics"
ora, prod = _full(depth_b200, which)
ora2 = copy.deepcopy(ora).cuda()
x, t = cases.full_batch()
x = x.to(torch.bfloat16).float()
ora.train(); ora2.train(); prod.train()
out_o = ora(x); ol.scale_invariant_loss(out_o.unsqueeze(1), t).backward()
with torch.autocast("cuda", dtype=torch.bfloat16):
    out_a = ora2(x.cuda())
ol.scale_invariant_loss(out_a.float().unsqueeze(1), t.cuda()).backward()
out_p = prod(x.cuda()); depth_b200.scale_invariant_loss(out_p.unsqueeze(1), t.cuda()).backward()
print("out: autocast", rel_err(out_a.detach().float().cpu(), out_o.detach()), " ours", rel_err(out_p.detach().cpu(), out_o.detach()))
go, ga = dict(ora.named_parameters()), dict(ora2.named_parameters())
rows = []
for k, p in prod.named_parameters():
    if k.startswith(("dinov2.", "pretrained.")) or go[k].grad is None or p.grad is None: continue
    if float(go[k].grad.norm()) < 1e-7: continue
    rows.append((k, rel_l2(ga[k].grad.float().cpu(), go[k].grad), rel_l2(p.grad.cpu(), go[k].grad)))
import statistics
for k, a, b in rows[::6]:
    print(f"  {k:58s} autocast {a:.3f}  ours {b:.3f}")
print("median autocast", statistics.median(r[1] for r in rows), "median ours", statistics.median(r[2] for r in rows))
print("max autocast", max(r[1] for r in rows), "max ours", max(r[2] for r in rows))
