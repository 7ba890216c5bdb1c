"""diagnostic: one eager step vs one graph replay from the same state; which parameters differ?"""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_b200 as pkg
from depth_b200 import ops
from oracle import fixtures as fx
from tests.test_benched_config_gpu import _pair, _batch, _opt

ora, prod = _pair(pkg)
eager = copy.deepcopy(prod)
cfg = fx.loss_config()
x, t = _batch(2, seed=100)
opt_e = _opt(eager)
w = eager.fusion_blocks[0].conv1.weight
print("version before", w._version)
for p in eager.parameters():
    p.grad = None
total, out = pkg.util.combined_loss_device(eager(x).unsqueeze(1), t, cfg, rgb=x)
total.backward()
ge = {k: p.grad.detach().clone() for k, p in eager.named_parameters() if p.grad is not None}
opt_e.step()
print("version after step", w._version)
opt_g = _opt(prod)
g = pkg.GraphedTrainStep(prod, opt_g, cfg, x, t, use_rgb=True, world=1, warmup=2)
g(x, t)
torch.cuda.synchronize()
gg = {k: p.grad.detach().clone() for k, p in prod.named_parameters() if p.grad is not None}
se, sg = eager.state_dict(), prod.state_dict()
bad = [(k, float((se[k].float() - sg[k].float()).abs().max())) for k in se if not torch.equal(se[k], sg[k])]
print("params differing after 1 step:", len(bad), "of", len(se))
for k, e in sorted(bad, key=lambda kv: -kv[1])[:15]:
    print("  ", k, e)
badg = [(k, float((ge[k].float() - gg[k].float()).abs().max() / (ge[k].abs().max() + 1e-30))) for k in ge if k in gg and not torch.equal(ge[k], gg[k])]
print("grads differing:", len(badg), "of", len(ge), "missing in graph:", [k for k in ge if k not in gg][:5], "missing in eager:", [k for k in gg if k not in ge][:5])
for k, e in sorted(badg, key=lambda kv: -kv[1])[:15]:
    print("  ", k, e)
# second eager forward vs fresh packs
x2, t2 = _batch(2, seed=101)
with torch.no_grad():
    a = eager(x2)
    ops.PACKS.store.clear()
    b = eager(x2)
print("eager second forward: cached == fresh packs:", torch.equal(a, b), float((a - b).abs().max()))
