"""spatial_reduction BatchNorm running statistics after N steps: oracle (fp32) vs product eager (fused trunk off)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_b200 as pkg
from oracle import fixtures as fx, losses as ol
from tests.test_benched_config_gpu import _pair, _batch, _opt
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
steps = int(os.environ.get("STEPS", "5"))
ora, prod = _pair(pkg)
prod.fused_encoder = False
cfg = fx.loss_config()
batches = [_batch(2, seed=100 + i) for i in range(steps)]
opt_o = torch.optim.AdamW([p for p in ora.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
opt_e = _opt(prod)
key = "cross_attention.spatial_reduction.1."
for i, (x, t) in enumerate(batches):
    opt_o.zero_grad(set_to_none=True)
    l = ol.scale_invariant_loss(ora(x).unsqueeze(1), t); l.backward(); opt_o.step()
    for p in prod.parameters():
        p.grad = None
    total, out = pkg.util.combined_loss_device(prod(x).unsqueeze(1), t, cfg, rgb=x)
    total.backward(); opt_e.step()
    bo, bp = dict(ora.named_buffers()), dict(prod.named_buffers())
    for nm in ("running_mean", "running_var"):
        a, b = bp[key + nm].float(), bo[key + nm].float()
        d = (a - b).abs()
        j = int(d.argmax())
        print(f"step {i} {nm}: max|b| {float(b.abs().max()):.5f} mean|b| {float(b.abs().mean()):.5f} max|a-b| {float(d.max()):.5f} at ch {j}: ours {float(a[j]):.5f} oracle {float(b[j]):.5f}")
    w_o = dict(ora.named_parameters())["cross_attention.spatial_reduction.0.weight"]
    w_p = dict(prod.named_parameters())["cross_attention.spatial_reduction.0.weight"]
    print(f"   conv weight drift max {float((w_o - w_p).abs().max()):.3e} of {float(w_o.abs().max()):.3e}; weight grad None? {w_p.grad is None}")
