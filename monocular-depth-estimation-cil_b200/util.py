"""Drop-in for the loss / metric half of the reference's ``src/util.py`` and ``main.combined_loss``.

Same names, argument meaning and error behaviour (``assert`` on shape mismatch) as the reference
(util.py:24-219, main.py:51-89); the arithmetic runs in the fused sm_100a reductions of
csrc/loss_metrics.cu through the C ABI.  Tensors must live on a CUDA device - there is no CPU path.
"""
import ctypes

import torch

from . import _lib as L


def _prep(x):
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


def _bhw(pred):
    """(B, rows, cols) view used by the flat per-sample reductions."""
    B = pred.shape[0]
    W = pred.shape[-1]
    return B, pred.numel() // (B * W), W


# Data-parallel training shards the batch by sample.  Two terms of the loss are NOT sums of per-sample terms: SiLog
# averages d and d^2 over the masked pixels of the WHOLE batch (reference util.py:118-121) and the edge-aware term
# normalises the RGB gradient magnitude with the min / max over the WHOLE batch (util.py:70).  With torch.distributed
# initialised those batch-global quantities are exchanged (one 3-double sum, one min/max all-reduce of the per-block
# partials), so an N-GPU step computes exactly the single-device loss of the global batch; the SiLog gradient is scaled
# by the world size because the gradient all-reduce averages over ranks.  Set to False for
# per-replica semantics (nn.DataParallel-style).
SYNC_BATCH_MOMENTS = True


def _world():
    import torch.distributed as dist
    if SYNC_BATCH_MOMENTS and dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


class _Pass:
    """One fused moments pass over (pred, target[, rgb])."""

    def __init__(self, pred, target, rgb, flags, eps, sync_silog=True, sync_edge=True):
        lib = L.lib()
        # the kernels index every buffer with pred's (B, H, W): a smaller target / rgb would be read out of bounds
        # where the reference raises a shape / broadcast error
        assert target.numel() == pred.numel() and target.shape[0] == pred.shape[0] \
            and target.shape[-2:] == pred.shape[-2:], \
            "Pred and target must have the same shape, got {} and {}".format(tuple(pred.shape), tuple(target.shape))
        if rgb is not None:
            assert rgb.dim() == 4 and rgb.shape[0] == pred.shape[0] and rgb.shape[1] == 3 \
                and rgb.shape[-2:] == pred.shape[-2:], \
                "rgb must be (B,3,H,W) matching pred {}, got {}".format(tuple(pred.shape), tuple(rgb.shape))
        self.p, self.t = _prep(pred), _prep(target)
        self.rgb = _prep(rgb) if rgb is not None else None
        self.B, self.H, self.W = _bhw(self.p)
        self.flags, self.eps = flags, float(eps)
        dev = self.p.device
        self.ws_bytes = lib.dp_depth_moments_workspace(self.B, self.H, self.W)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.mom = torch.empty(self.B, L.NMOM, dtype=torch.float64, device=dev)
        self.mm = None
        if flags & L.F_EDGE:
            self.mm = torch.empty(lib.dp_rgb_minmax_bytes() // 4, dtype=torch.float32, device=dev)
            L.check(lib.dp_rgb_gradmag_minmax(L.ptr(self.rgb), self.B, self.H, self.W, L.ptr(self.mm), L.stream()))
            if sync_edge and _world() > 1:      # batch-global min / max (util.py:70): the partials are (min, max) pairs
                import torch.distributed as dist
                t = self.mm.view(-1, 2).clone()      # device-only arithmetic: the step is captured in a CUDA graph
                t[:, 1].neg_()
                dist.all_reduce(t, op=dist.ReduceOp.MIN)        # min of mins, min of negated maxes
                t[:, 1].neg_()
                self.mm.copy_(t.view(-1))
        L.check(lib.dp_depth_moments(L.ptr(self.p), L.ptr(self.t), L.ptr(self.rgb), L.ptr(self.mm), self.B, self.H,
                                     self.W, flags, self.eps, L.ptr(self.mom), L.ptr(self.ws), self.ws_bytes,
                                     L.stream()))
        self.world = _world() if sync_silog else 1
        if self.world > 1 and (flags & L.F_SILOG):
            # batch-global SiLog moments (count, sum d, sum d^2 over the masked pixels of every rank's samples): the
            # combine / backward kernels add the per-sample rows, so the global sums go into row 0 and the rest is zero
            import torch.distributed as dist
            tot = self.mom[:, L.M_M0:L.M_M2 + 1].sum(dim=0)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            self.mom[:, L.M_M0:L.M_M2 + 1] = 0
            self.mom[0, L.M_M0:L.M_M2 + 1] = tot

    def combine(self, w_si=1.0, w_silog=0.0, vf=0.85, w_grad=0.0, beta=0.0, sqroot=False, per_sample=False):
        out = torch.empty(L.NLOSS, dtype=torch.float32, device=self.p.device)
        ps = torch.empty(self.B, dtype=torch.float32, device=self.p.device) if per_sample else None
        L.check(L.lib().dp_loss_combine(L.ptr(self.mom), self.B, self.H, self.W, self.flags, w_si, w_silog, vf, w_grad,
                                        beta, int(sqroot), L.ptr(out), L.ptr(ps), L.stream()))
        return out, ps

    def backward(self, grad_out, w_si, w_silog, vf, w_grad, beta, si_scale=None):
        g = torch.empty_like(self.p)
        go = grad_out.detach().float().contiguous() if grad_out is not None else None
        if self.world > 1:
            w_silog = w_silog * self.world      # the gradient all-reduce averages over ranks; SiLog is already global
        L.check(L.lib().dp_loss_backward(L.ptr(self.p), L.ptr(self.t), L.ptr(self.rgb), L.ptr(self.mm), L.ptr(self.mom),
                                         L.ptr(go), L.ptr(si_scale), self.B, self.H, self.W, self.flags, self.eps, w_si,
                                         w_silog, vf, w_grad, beta, L.ptr(g), L.stream()))
        return g

    def counts(self, thresholds, aligned=True, eps_div=0.0):
        n = len(thresholds)
        cnt = torch.empty(self.B, n, dtype=torch.int64, device=self.p.device)
        arr = (ctypes.c_float * n)(*[float(t) for t in thresholds])
        L.check(L.lib().dp_delta_counts(L.ptr(self.p), L.ptr(self.t), L.ptr(self.mom), self.B, self.H, self.W, arr, n,
                                        int(aligned), float(eps_div), L.ptr(cnt), L.ptr(self.ws), self.ws_bytes,
                                        L.stream()))
        return cnt


class _LossFn(torch.autograd.Function):
    """total = w_si*SI + w_silog*SiLog + w_grad*Grad + Edge(beta); returns (out[8] scalars)."""

    @staticmethod
    def forward(ctx, pred, target, rgb, flags, eps, w_si, w_silog, vf, w_grad, beta, sqroot, slot):
        # the batch-global exchanges are made for the terms that carry weight; a zero-weight term is only logged
        # (main.py:85-88) and then reports this rank's shard
        ps = _Pass(pred, target, rgb, flags, eps, sync_silog=w_silog != 0.0, sync_edge=beta != 0.0)
        out, per = ps.combine(w_si, w_silog, vf, w_grad, beta, sqroot, per_sample=bool(sqroot))
        ctx.ps, ctx.per, ctx.args, ctx.slot = ps, per, (w_si, w_silog, vf, w_grad, beta, sqroot), slot
        ctx.in_shape, ctx.in_dtype = pred.shape, pred.dtype
        ctx.mark_non_differentiable(out)
        return out[slot].clone(), out

    @staticmethod
    def backward(ctx, g_val, _g_out):
        w_si, w_silog, vf, w_grad, beta, sqroot = ctx.args
        if ctx.slot != L.L_TOTAL:
            # a single weighted term was requested: switch the others off
            w_si = w_si if ctx.slot in (L.L_SI, L.L_SI_RAW) else 0.0
            w_silog = w_silog if ctx.slot in (L.L_SILOG, L.L_SILOG_RAW) else 0.0
            w_grad = w_grad if ctx.slot == L.L_GRAD else 0.0
            beta = beta if ctx.slot == L.L_EDGE else 0.0
        scale = None
        if sqroot:
            scale = 0.5 / ctx.per
        g = ctx.ps.backward(g_val, w_si, w_silog, vf, w_grad, beta, scale)
        ctx.ps = None
        return (g.reshape(ctx.in_shape).to(ctx.in_dtype),) + (None,) * 11


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.DepthB200Error("depth_b200.util functions run on CUDA tensors only (no CPU fallback)")


def scale_invariant_loss(pred, target, epsilon=1e-6, sqroot=False):
    """reference util.py:129-156."""
    assert pred.shape[-2:] == target.shape[-2:], \
        "Pred and target must have the same spatial dimensions, got {} and {}".format(pred.shape[-2:], target.shape[-2:])
    _check_cuda(pred, target)
    val, _ = _LossFn.apply(pred, target, None, L.F_SI, epsilon, 1.0, 0.0, 0.85, 0.0, 0.0, bool(sqroot), L.L_SI_RAW)
    return val


def silog_loss(pred, target, mask=None, variance_focus=0.85, epsilon=1e-6):
    """reference util.py:90-127.  mask=None or the reference's own call-site mask (target > 0) (main.py:69)."""
    _check_cuda(pred, target)
    if pred.shape != target.shape:
        target = torch.nn.functional.interpolate(target, size=pred.shape[2:], mode="bilinear", align_corners=True)
    if mask is not None:
        # the fused pass masks with (target > 0); any other mask is folded into target (masked-out -> 0 -> excluded)
        if mask.dtype != torch.bool:
            mask = mask.bool()
        if not torch.equal(mask, target > 0):
            assert bool((target[mask] > 0).all()), "explicit masks must select positive targets"
            target = torch.where(mask, target, torch.zeros_like(target))
    val, _ = _LossFn.apply(pred, target, None, L.F_SI | L.F_SILOG, epsilon, 0.0, 1.0, float(variance_focus), 0.0, 0.0,
                           False, L.L_SILOG)
    return val


def gradient_loss(pred, target):
    """reference util.py:24-44."""
    _check_cuda(pred, target)
    assert pred.shape == target.shape and pred.dim() == 4 and pred.shape[1] == 1
    val, _ = _LossFn.apply(pred, target, None, L.F_GRAD, 1e-6, 0.0, 0.0, 0.85, 1.0, 0.0, False, L.L_GRAD)
    return val


def edge_aware_loss(pred, target, rgb, beta=0.5):
    """reference util.py:46-88."""
    _check_cuda(pred, target, rgb)
    assert pred.shape == target.shape and pred.dim() == 4 and pred.shape[1] == 1 and rgb.shape[1] == 3
    val, _ = _LossFn.apply(pred, target, rgb, L.F_GRAD | L.F_EDGE, 1e-6, 0.0, 0.0, 0.85, 0.0, float(beta), False,
                           L.L_EDGE)
    return val


def combined_loss_device(pred, target, config, rgb=None):
    """combined_loss without the host read: returns (total, device tensor of the DP_NLOSS scalars).  Capturable in a
    CUDA graph (no synchronisation)."""
    _check_cuda(pred, target, rgb)
    assert pred.shape[-2:] == target.shape[-2:], \
        "Pred and target must have the same spatial dimensions, got {} and {}".format(pred.shape[-2:], target.shape[-2:])
    lf = config.model.loss_function
    flags = L.F_SI | L.F_SILOG | L.F_GRAD | (L.F_EDGE if rgb is not None else 0)
    return _LossFn.apply(pred, target, rgb, flags, 1e-6, float(lf.si_loss_alpha), float(lf.silog_loss.alpha),
                         float(lf.silog_loss.variance_focus), float(lf.grad_loss_alpha),
                         float(lf.edge_loss_alpha) if rgb is not None else 0.0, False, L.L_TOTAL)


def combined_loss(pred, target, config, rgb=None):
    """reference main.py:51-89: returns (total, {'si_loss','silog_loss','grad_loss','edge_loss'}) - one fused
    pass and ONE device->host read instead of four `.item()` syncs."""
    _check_cuda(pred, target, rgb)
    assert pred.shape[-2:] == target.shape[-2:], \
        "Pred and target must have the same spatial dimensions, got {} and {}".format(pred.shape[-2:], target.shape[-2:])
    lf = config.model.loss_function
    flags = L.F_SI | L.F_SILOG | L.F_GRAD | (L.F_EDGE if rgb is not None else 0)
    total, out = _LossFn.apply(pred, target, rgb, flags, 1e-6, float(lf.si_loss_alpha), float(lf.silog_loss.alpha),
                               float(lf.silog_loss.variance_focus), float(lf.grad_loss_alpha),
                               float(lf.edge_loss_alpha) if rgb is not None else 0.0, False, L.L_TOTAL)
    host = out.tolist()
    return total, {"si_loss": host[L.L_SI], "silog_loss": host[L.L_SILOG], "grad_loss": host[L.L_GRAD],
                   "edge_loss": host[L.L_EDGE] if rgb is not None else 0.0}


def absolute_relative_error(pred, target):
    """reference util.py:210-219."""
    assert pred.shape == target.shape, \
        "Pred and target must have the same shape, got {} and {}".format(pred.shape, target.shape)
    _check_cuda(pred, target)
    ps = _Pass(pred, target, None, L.F_ABSREL, 1e-6)
    out, _ = ps.combine()
    return out[L.L_ABSREL]


def delta_thres(pred, target, thres=0.1):
    """reference util.py:183-207."""
    assert pred.shape == target.shape, \
        "Pred and target must have the same shape, got {} and {}".format(pred.shape, target.shape)
    _check_cuda(pred, target)
    return evaluation_metrics(pred, target, [thres])[2]


def delta_counts(pred, target, thresholds, aligned=True):
    """integer per-sample pixel counts behind delta_thres (B, len(thresholds))."""
    _check_cuda(pred, target)
    ps = _Pass(pred, target, None, L.F_SI, 1e-6)
    return ps.counts(thresholds, aligned=aligned, eps_div=0.0 if aligned else 1e-6)


def evaluation_metrics(pred, target, thresholds=(1.05, 1.05 ** 2, 1.05 ** 3), fast_math=None):
    """The metric set of evaluation.py:157-166 for one batch in one streaming kernel (each input read from HBM once and
    classified from shared memory):
    returns a device tensor [SI-RMSE, AbsRel, delta_1 .. delta_k] (batch means, as the reference's functions).
    fast_math=None (default): the lean arithmetic (one shared reciprocal + one lg2 per pixel, division-free threshold
    test; exact code for slices with negative values) - within the 1e-5 relative / 0.01 %-of-pixels contract by a wide
    margin (measured ~1e-7 / a few ppm).  fast_math=False: IEEE logf / division, the reference's own arithmetic (the
    checker the other modes are tested against).  fast_math=True: MUFU lg2 / rcp per operand (round-1 variant)."""
    assert pred.shape == target.shape, \
        "Pred and target must have the same shape, got {} and {}".format(pred.shape, target.shape)
    _check_cuda(pred, target)
    p, t = _prep(pred), _prep(target)
    B, H, W = _bhw(p)
    n = len(thresholds)
    dev = p.device
    mom = torch.empty(B, L.NMOM, dtype=torch.float64, device=dev)
    cnt = torch.empty(B, n, dtype=torch.int64, device=dev)
    out = torch.empty(2 + n, dtype=torch.float32, device=dev)
    arr = (ctypes.c_float * n)(*[float(x) for x in thresholds])
    ws = torch.empty(L.lib().dp_eval_metrics_workspace(B, H, W), dtype=torch.uint8, device=dev)
    mode = 2 if fast_math is None else int(bool(fast_math))
    L.check(L.lib().dp_eval_metrics(L.ptr(p), L.ptr(t), B, H, W, arr, n, 1e-6, mode, L.ptr(mom), L.ptr(cnt),
                                    L.ptr(out), L.ptr(ws), ws.numel(), L.stream()))
    return out


def evaluate_model_sums(outputs, targets):
    """Raw sums of main.evaluate_model's loop body (main.py:291-321) for one batch, as a dict of Python numbers:
    abs, sq, rel, sirmse (sum over images), d1..d3 (unaligned delta counts at 1.25^k)."""
    _check_cuda(outputs, targets)
    if outputs.shape[-2:] != targets.shape[-2:]:
        outputs = torch.nn.functional.interpolate(outputs, size=targets.shape[-2:], mode="bilinear", align_corners=True)
    ps = _Pass(outputs, targets, None, L.F_M4 | L.F_ABSREL, 1e-6)
    cnt = ps.counts([1.25, 1.25 ** 2, 1.25 ** 3], aligned=False, eps_div=1e-6)
    m = ps.mom.cpu()
    c = cnt.cpu()
    v0, v1, v2 = m[:, L.M_V0], m[:, L.M_V1], m[:, L.M_V2]
    var = (v2 / v0 - (v1 / v0) ** 2).clamp_min(0)
    sir = torch.where(v0 > 0, var.sqrt(), torch.zeros_like(var)).sum().item()
    return {"abs": m[:, L.M_AB].sum().item(), "sq": m[:, L.M_SQ].sum().item(), "rel": m[:, L.M_AR].sum().item(),
            "sirmse": sir, "d1": int(c[:, 0].sum()), "d2": int(c[:, 1].sum()), "d3": int(c[:, 2].sum())}


def per_pixel_scale_invariant_loss(pred, target):
    """reference util.py:159-181 (visualisation; single image (H,W) or any (..., H, W) stack treated image by image):
    (d - mean d)^2 with d = log pred - log target, no epsilon.  One moments pass for the per-image mean and one
    elementwise kernel (dp_per_pixel_si)."""
    assert pred.shape == target.shape, \
        "Pred and target must have the same shape, got {} and {}".format(pred.shape, target.shape)
    _check_cuda(pred, target)
    assert (pred > 0).all() and (target > 0).all(), "Pred and target must be positive"
    p, t = _prep(pred), _prep(target)
    H, W = p.shape[-2], p.shape[-1]
    B = p.numel() // (H * W)
    ps = _Pass(p.reshape(B, 1, H, W), t.reshape(B, 1, H, W), None, L.F_SI, 0.0)
    out = torch.empty_like(p)
    L.check(L.lib().dp_per_pixel_si(L.ptr(p), L.ptr(t), L.ptr(ps.mom), B, H, W, L.ptr(out), L.stream()))
    return out


def remove_module_prefix(state_dict):
    """reference util.py:14-22."""
    from collections import OrderedDict
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k.replace("module.", "", 1) if k.startswith("module.") else k] = v
    return out


def load_model(model_type, checkpoint_path, model_cfg=None):
    """reference util.py:222-238 / evaluation.py:42-66: build the configured model and load a checkpoint written by
    main.py (``{'model_state_dict': ...}``) or a bare / ``module.``-prefixed state_dict."""
    from .network.midas_net_custom import MidasNet_small
    from .network.midas_semantics import MidasNetSemantics
    if model_type != 'MiDaS_small':
        raise NotImplementedError(f"model_type {model_type!r}: the reference's load_model only builds 'MiDaS_small'")
    if getattr(model_cfg, "dinov2_type", None) is not None:
        model = MidasNetSemantics(None, features=64, backbone="efficientnet_lite3", exportable=True, non_negative=True,
                                  cfg=model_cfg, blocks={'expand': True}, dinov2_type=model_cfg.dinov2_type)
    else:
        model = MidasNet_small(None, features=64, backbone="efficientnet_lite3", exportable=True, non_negative=True,
                               cfg=model_cfg, blocks={'expand': True})
    checkpoint = torch.load(checkpoint_path, map_location="cpu")
    if 'model_state_dict' in checkpoint:
        model.load_state_dict(checkpoint['model_state_dict'])
    else:
        model.load_state_dict(remove_module_prefix(checkpoint))
    return model


def ensure_dir(directory):
    """reference util.py:288-290."""
    import os
    if not os.path.exists(directory):
        os.makedirs(directory)


def generate_test_predictions(model, test_loader, device, predictions_dir, size=(426, 560)):
    """reference util.py:292-325: eval-mode forward, bilinear resize of the depth maps to the dataset's native 426x560
    (align_corners=True) and one ``<name>.npy`` per sample.  The resize runs in the fp32 plane kernel
    (dp_resize_bilinear_planes_f32) and each batch leaves the device in ONE asynchronous copy into pinned memory instead
    of a ``.cpu()`` per sample; file names follow the reference (second token of the list entry)."""
    import os
    import numpy as np
    from . import ops
    model.eval()
    ensure_dir(predictions_dir)
    host = None
    with torch.no_grad():
        for inputs, filenames in test_loader:
            inputs = inputs.to(device, non_blocking=True)
            outputs = model(inputs).unsqueeze(1)                               # (B,1,H,W) fp32
            outputs = ops.resize_planes_f32(outputs, size, True)               # (B,1,426,560)
            if host is None or host.shape != outputs.shape:
                host = torch.empty(outputs.shape, dtype=torch.float32).pin_memory()
            host.copy_(outputs, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            arr = host.numpy()
            for i in range(arr.shape[0]):
                filename = filenames[i].split(' ')[1]
                np.save(os.path.join(predictions_dir, f"{filename}"), arr[i, 0])
