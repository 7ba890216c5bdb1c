"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the reference's loss and metric arithmetic, written as plain torch tensor
algebra so it can be evaluated in fp32 (what the reference does) or fp64 (to bound rounding).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Pinned against the reference itself: oracle/make_golden.py imports /root/reference/src/util.py in
the build container, checks every function below bit-for-bit in fp32 against it on seeded inputs,
and stores the reference's outputs in tests/golden/loss_metric_golden.json.

Each function cites the reference lines it follows.
"""
import torch
import torch.nn.functional as F


def _logdiff(pred, target, eps):
    return torch.log(pred + eps) - torch.log(target + eps)


def scale_invariant_loss(pred, target, epsilon=1e-6, sqroot=False):
    """reference src/util.py:129-156: per-sample E[d^2] - E[d]^2 of d = log(p+eps)-log(t+eps); batch mean."""
    assert pred.shape[-2:] == target.shape[-2:]
    d = _logdiff(pred, target, epsilon)
    n = d.numel() / d.shape[0]
    per = (d ** 2).sum(dim=[1, 2, 3]) / n - d.sum(dim=[1, 2, 3]) ** 2 / (n ** 2)
    if sqroot:
        per = per.sqrt()
    return per.mean()


def silog_loss(pred, target, mask=None, variance_focus=0.85, epsilon=1e-6):
    """reference src/util.py:90-127: moments over all masked pixels of the whole batch."""
    if pred.shape != target.shape:
        target = F.interpolate(target, size=pred.shape[2:], mode="bilinear", align_corners=True)
    if mask is None:
        mask = target > 0
    d = _logdiff(pred[mask], target[mask], epsilon)
    return (d ** 2).mean() - variance_focus * d.mean() ** 2


def _absdx(x):
    return (x[..., :, :-1] - x[..., :, 1:]).abs()


def _absdy(x):
    return (x[..., :-1, :] - x[..., 1:, :]).abs()


def gradient_loss(pred, target):
    """reference src/util.py:24-44."""
    return (_absdx(pred) - _absdx(target)).abs().mean() + (_absdy(pred) - _absdy(target)).abs().mean()


def edge_aware_loss(pred, target, rgb, beta=0.5):
    """reference src/util.py:46-88: RGB gradient magnitude, globally min/max normalised, weights the
    zero-padded depth-gradient differences; means are over the full (padded) HxW grid."""
    gx = F.pad(_absdx(rgb), (0, 1, 0, 0))
    gy = F.pad(_absdy(rgb), (0, 0, 0, 1))
    g = torch.sqrt(gx.pow(2).mean(dim=1, keepdim=True) + gy.pow(2).mean(dim=1, keepdim=True))
    g = (g - g.min()) / (g.max() - g.min() + 1e-6)
    ex = F.pad(_absdx(pred), (0, 1, 0, 0)) - F.pad(_absdx(target), (0, 1, 0, 0))
    ey = F.pad(_absdy(pred), (0, 0, 0, 1)) - F.pad(_absdy(target), (0, 0, 0, 1))
    return beta * ((g * ex.abs()).mean() + (g * ey.abs()).mean())


def combined_loss(pred, target, config, rgb=None):
    """reference src/main.py:51-89; config exposes model.loss_function.{si_loss_alpha,silog_loss.{alpha,
    variance_focus},grad_loss_alpha,edge_loss_alpha}."""
    lf = config.model.loss_function
    si = scale_invariant_loss(pred, target) * lf.si_loss_alpha
    sl = silog_loss(pred, target, mask=(target > 0).detach(), variance_focus=lf.silog_loss.variance_focus) \
        * lf.silog_loss.alpha
    gr = gradient_loss(pred, target) * lf.grad_loss_alpha
    ed = 0.0
    if rgb is not None:
        ed = edge_aware_loss(pred, target, rgb, lf.edge_loss_alpha)
    total = si + sl + gr + ed
    return total, {"si_loss": si.item(), "silog_loss": sl.item(), "grad_loss": gr.item(),
                   "edge_loss": ed.item() if rgb is not None else 0.0}


def absolute_relative_error(pred, target):
    """reference src/util.py:210-219."""
    assert pred.shape == target.shape
    return ((target - pred).abs() / (target + 1e-6)).mean()


def delta_thres(pred, target, thres=0.1):
    """reference src/util.py:183-207: per-sample log-mean scale alignment, then max(a/t, t/a) < thres."""
    assert pred.shape == target.shape
    eps = 1e-6
    B = pred.shape[0]
    p = pred.reshape(B, -1)
    t = target.reshape(B, -1)
    s = torch.exp((torch.log(t + eps) - torch.log(p + eps)).mean(dim=1, keepdim=True))
    a = p * s
    ratio = torch.max(a / t, t / a)
    return (ratio < thres).float().mean(dim=1).mean()


def delta_counts(pred, target, thresholds):
    """Integer per-sample pixel counts behind delta_thres (same arithmetic), for the 0.01 %-of-pixels gate."""
    eps = 1e-6
    B = pred.shape[0]
    p = pred.reshape(B, -1)
    t = target.reshape(B, -1)
    s = torch.exp((torch.log(t + eps) - torch.log(p + eps)).mean(dim=1, keepdim=True))
    a = p * s
    ratio = torch.max(a / t, t / a)
    return torch.stack([(ratio < th).sum(dim=1) for th in thresholds], dim=1)


def per_pixel_scale_invariant_loss(pred, target):
    """reference src/util.py:159-181 (single image, no eps)."""
    assert pred.shape == target.shape
    assert (pred > 0).all() and (target > 0).all()
    d = torch.log(pred) - torch.log(target)
    return (d - d.mean()) ** 2


def evaluate_metric_sums(outputs, targets):
    """reference src/main.py:254-392 metric set for ONE batch: returns the raw sums the loop accumulates
    (sum|p-t|, sum(p-t)^2, sum|p-t|/(t+1e-6), sum of per-image siRMSE over t>1e-6 with p clamped to 1e-6,
    unaligned delta counts at 1.25^k).  outputs are resized to the target size first (align_corners=True)."""
    if outputs.shape[-2:] != targets.shape[-2:]:
        outputs = F.interpolate(outputs, size=targets.shape[-2:], mode="bilinear", align_corners=True)
    ad = (outputs - targets).abs()
    res = {"abs": ad.sum().item(), "sq": (ad ** 2).sum().item(), "rel": (ad / (targets + 1e-6)).sum().item()}
    sir = 0.0
    for i in range(outputs.shape[0]):
        p = outputs[i].reshape(-1).double()
        t = targets[i].reshape(-1).double()
        v = t > 1e-6
        if not bool(v.any()):
            continue
        lp = torch.log(torch.where(p[v] > 1e-6, p[v], torch.full_like(p[v], 1e-6)).float()).double()
        lt = torch.log(t[v].float()).double()
        d = (lp.float() - lt.float())
        sir += float(torch.sqrt(((d - d.mean()) ** 2).mean()))
    res["sirmse"] = sir
    r = torch.max(outputs / (targets + 1e-6), targets / (outputs + 1e-6))
    for k in (1, 2, 3):
        res[f"d{k}"] = int((r < 1.25 ** k).sum().item())
    return res


def evaluate_model(model, val_loader, device="cpu"):
    """reference src/main.py:254-392 `evaluate_model`: the whole loop (eval mode, no grad, resize to the target size,
    per-batch sums, per-image numpy siRMSE, final normalisation by N * C * H * W) -> dict MAE, RMSE, siRMSE, REL,
    Delta1..3.  Pinned against the reference's own function by oracle/make_golden.py (identity model, two batches,
    prediction resolution != target resolution)."""
    import math
    model.eval()
    acc = {"abs": 0.0, "sq": 0.0, "rel": 0.0, "sirmse": 0.0, "d1": 0.0, "d2": 0.0, "d3": 0.0}
    total_samples, target_shape = 0, None
    with torch.no_grad():
        for inputs, targets, _names in val_loader:
            inputs, targets = inputs.to(device), targets.to(device)
            total_samples += inputs.size(0)
            if target_shape is None:
                target_shape = targets.shape
            outputs = model(inputs).unsqueeze(1)
            sums = evaluate_metric_sums(outputs, targets)
            for k in acc:
                acc[k] += sums[k]
    total_pixels = target_shape[1] * target_shape[2] * target_shape[3]
    n = total_samples * total_pixels
    return {"MAE": acc["abs"] / n, "RMSE": math.sqrt(acc["sq"] / n), "siRMSE": acc["sirmse"] / total_samples,
            "REL": acc["rel"] / n, "Delta1": acc["d1"] / n, "Delta2": acc["d2"] / n, "Delta3": acc["d3"] / n}
