import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_b200
from oracle import cases, fixtures as fx, losses as ol
from tests.test_modules_gpu import _full, rel_err, rel_l2
which = sys.argv[1] if len(sys.argv) > 1 else "small"
ora, prod = _full(depth_b200, which)
x, t = cases.full_batch()
x = x.to(torch.bfloat16).float()
ora.train(); prod.train()
out_o = ora(x); loss_o = ol.scale_invariant_loss(out_o.unsqueeze(1), t); loss_o.backward()
out_p = prod(x.cuda()); loss_p = depth_b200.scale_invariant_loss(out_p.unsqueeze(1), t.cuda()); loss_p.backward()
print("out", rel_err(out_p.detach().cpu(), out_o.detach()), rel_l2(out_p.detach().cpu(), out_o.detach()), "loss", loss_p.item(), loss_o.item())
print("zero frac", float((out_o == 0).float().mean()), float((out_p == 0).float().mean()))
go = dict(ora.named_parameters())
for k, p in prod.named_parameters():
    if k.startswith("dinov2."): continue
    if k.startswith("pretrained.") and not k.endswith("0.weight"): continue
    if go[k].grad is None:
        print(f"   {k:60s} ref None, prod {'None' if p.grad is None else 'SET'}"); continue
    if p.grad is None:
        print(f"   {k:60s} prod None!"); continue
    print(f"   {k:60s} l2 {rel_l2(p.grad.cpu(), go[k].grad):.4f}  norm ref {float(go[k].grad.norm()):.3e} prod {float(p.grad.norm()):.3e}")
