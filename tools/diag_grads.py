"""diagnostic: parameter-gradient drift from the fp32 oracle at 448x576, ours vs stock bf16 autocast, by module group"""
import copy, sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
import depth_b200 as pkg
from oracle import fixtures as fx, losses as ol
from tests.test_benched_config_gpu import _pair, _batch, rel_l2, rel_max

mode = sys.argv[1] if len(sys.argv) > 1 else "train_si"
ora, prod = _pair(pkg)
prod.fused_encoder = False
auto = copy.deepcopy(ora)
x, t = _batch(4)
if mode.startswith("eval"):
    ora.eval(); prod.eval(); auto.eval()
w = (torch.rand(4, 448, 576, generator=torch.Generator().manual_seed(5)) + 0.5).cuda()

def loss_fn(out, which):
    if mode.endswith("si"):
        return ol.scale_invariant_loss(out.unsqueeze(1), t) if which != "prod" else pkg.scale_invariant_loss(out.unsqueeze(1), t)
    return (out * w).mean()

loss_fn(ora(x), "ora").backward()
loss_fn(prod(x), "prod").backward()
with torch.autocast("cuda", dtype=torch.bfloat16):
    oa = auto(x)
loss_fn(oa.float(), "auto").backward()
go, ga = dict(ora.named_parameters()), dict(auto.named_parameters())
groups = collections.defaultdict(lambda: ([], []))
gmax = max(float(p.grad.norm()) for p in go.values() if p.grad is not None)
for k, p in prod.named_parameters():
    if k.startswith("dinov2.") or go[k].grad is None or p.grad is None:
        continue
    if float(go[k].grad.norm()) < 1e-6 * gmax:
        continue
    grp = ".".join(k.split(".")[:2])
    groups[grp][0].append(rel_l2(p.grad, go[k].grad))
    groups[grp][1].append(rel_l2(ga[k].grad.float(), go[k].grad))
print(mode)
for g, (o, s) in groups.items():
    print(f"  {g:40s} n={len(o):3d} ours med {np.median(o):.4f} max {np.max(o):.4f} | autocast med {np.median(s):.4f} max {np.max(s):.4f}")
