"""Autograd-aware host wrappers over the C ABI (include/depth_b200.h).

Activations on this path are NHWC bf16 torch tensors of shape (B, H, W, C) whose last dim is contiguous
(a channel slice of a wider buffer is allowed: the pixel stride `ld` is taken from the tensor strides).
Every function here launches hand-written sm_100a kernels; torch only owns memory, streams and the
autograd tape.  Nothing falls back to ATen for the arithmetic.
"""
import weakref

import torch

from . import _lib as L

BF16 = torch.bfloat16


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def _ld(x):
    """pixel stride (elements) of an NHWC view; validates the layout."""
    B, H, W, C = x.shape
    assert x.dtype == BF16 and x.is_cuda, "NHWC bf16 CUDA tensor expected"
    assert x.stride(3) == 1, "channels must be contiguous"
    ld = x.stride(2) if W > 1 else (x.stride(1) if H > 1 else max(x.stride(0), C))
    if W > 1 and H > 1:
        assert x.stride(1) == W * ld, "rows must be densely packed"
    if B > 1 and H * W > 1:
        assert x.stride(0) == H * W * ld, "images must be densely packed"
    assert ld % 8 == 0 and x.data_ptr() % 16 == 0, "pixel stride / base must be 16-byte aligned"
    return ld


def _nhwc(B, H, W, C, dev):
    return torch.empty(B, H, W, C, dtype=BF16, device=dev)


def mark_internal(t):
    """Tag a tensor (or each tensor of a tuple) as being in the internal NHWC bf16 layout.  Module boundaries
    (network.blocks.enter) trust this explicit tag, never the dtype: a caller's own bf16 NCHW tensor is a public tensor
    and is converted like an fp32 one."""
    if isinstance(t, tuple):
        for u in t:
            mark_internal(u)
    elif torch.is_tensor(t) and t.dtype == BF16 and t.dim() == 4:
        t._dp_nhwc = True
    return t


def is_internal(t):
    return bool(getattr(t, "_dp_nhwc", False))


def _internal(fn):
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        return mark_internal(fn(*a, **k))
    return wrapped


def _dense(t):
    """gradient tensors arrive with arbitrary strides: make them dense NHWC bf16."""
    if t is None:
        return None
    if t.dtype != BF16:
        t = t.to(BF16)
    return t if t.is_contiguous() else t.contiguous()


class _PackCache:
    """bf16 kernel-layout copies of fp32 parameters, refreshed when the parameter changes in place.

    `_version` only sees updates made through the dispatcher.  A CUDA-graph replay (graphs.GraphedTrainStep) updates the
    parameters on the device without touching it, so every replay also bumps `generation`, which is part of the
    freshness key: the first eager use of a weight after any number of replays repacks it."""

    def __init__(self):
        self.store = {}
        self.generation = 0

    def invalidate(self):
        """parameters changed behind autograd's back (graph replay, raw device writes)"""
        self.generation += 1

    def _ver(self, w):
        return (w._version, w.data_ptr(), tuple(w.shape), self.generation)

    def get(self, w, swap, flip):
        key = (id(w), swap, flip)
        ver = self._ver(w)
        hit = self.store.get(key)
        if hit is not None and hit[0] == ver and hit[2]() is w:
            return hit[1]
        D0, D1, KH, KW = w.shape
        A, Bc = (D1, D0) if swap else (D0, D1)
        # the pack buffer of a (weight, layout) pair is kept across refreshes: stable addresses let repack_all() refresh
        # every pack of a step with one launch from a cached descriptor table
        if hit is not None and hit[2]() is w and hit[1].shape == (KH * KW, A, Bc) and hit[1].device == w.device:
            dst = hit[1]
        else:
            dst = torch.empty(KH * KW, A, Bc, dtype=BF16, device=w.device)
            self._table = None
        src = w.detach()
        if src.dtype != torch.float32 or not src.is_contiguous():
            src = src.float().contiguous()
        L.check(L.lib().dp_pack_conv_weight(L.ptr(src), D0, D1, KH, KW, int(swap), int(flip), L.ptr(dst), Bc, L.stream()))
        self.store[key] = (ver, dst, weakref.ref(w))
        if len(self.store) > 4096:
            self.store.clear()
            self._table = None
        return dst

    _table = None

    def repack_all(self):
        """refresh every known pack (all weights a previous step asked for, in the layouts it asked for) with ONE launch;
        the per-weight get() calls of the step that follows then hit the cache.  Used at the head of the captured train
        step (graphs.GraphedTrainStep): ~230 tiny pack kernels become one."""
        import struct
        items = []
        for key, (ver, dst, ref) in list(self.store.items()):
            w = ref()
            # parameters only: temporaries (e.g. the re-viewed weight of a non-overlapping transposed conv) come and go
            if w is None or not w.is_leaf or not w.is_cuda or w.dtype != torch.float32 or not w.is_contiguous() \
                    or w.dim() != 4:
                continue
            items.append((key, w, dst))
        if not items:
            return 0
        sig = tuple((k, w.data_ptr(), d.data_ptr()) for k, w, d in items)
        if self._table is None or self._table[0] != sig:
            if torch.cuda.is_current_stream_capturing():
                return 0          # the set of packs changed under capture: fall back to the per-weight packs
            buf = bytearray()
            for (wid, swap, flip), w, dst in items:
                D0, D1, KH, KW = w.shape
                buf += struct.pack("<QQ8i", w.data_ptr(), dst.data_ptr(), D0, D1, KH, KW, int(swap), int(flip), dst.shape[2], 0)
            host = torch.frombuffer(buf, dtype=torch.uint8).clone().pin_memory()
            self._table = (sig, host.to(items[0][1].device, non_blocking=False))
        L.check(L.lib().dp_pack_conv_weights_batched(L.ptr(self._table[1]), len(items), 8, L.stream()))
        for key, w, dst in items:
            self.store[key] = (self._ver(w), dst, weakref.ref(w))
        return len(items)


PACKS = _PackCache()

# Fused / capturable optimizers (torch.optim.AdamW(fused=True)) update parameters WITHOUT bumping `_version`
# (measured: `_version` stays 0 across AdamW(fused=True).step()), so an eager training loop would keep convolving with
# the packs of step 0.  Every optimizer step, of any optimizer, therefore invalidates the cache.
try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook
    _reg_post_hook(lambda _opt, _args, _kwargs: PACKS.invalidate())
except ImportError:  # pragma: no cover - very old torch: fall back to never trusting the cache across calls
    _PackCache.get = (lambda orig: (lambda self, w, swap, flip: (self.invalidate(), orig(self, w, swap, flip))[1]))(
        _PackCache.get)


# --------------------------------------------------------------------------------------------------
# weight gradients on a side stream (opt-in: graphs.GraphedTrainStep)
# --------------------------------------------------------------------------------------------------
class _Side:
    """Weight gradients feed nothing but the optimizer, so inside a captured train step they are issued on a second
    stream and overlap the (mostly bandwidth-bound) kernels of the data-gradient chain.  The operands are kept alive
    until the next join so the caching allocator cannot hand their memory to the main stream while the side stream
    still reads it; joins happen when more than `limit` bytes are pending and once after backward."""
    enabled = False
    stream = None
    pending = []
    pending_bytes = 0
    limit = 6 << 30
    reducer = None        # distributed.GradientAllReducer of the step being run (graphs.GraphedTrainStep sets it)


def side_enable(flag):
    _Side.enabled = bool(flag)
    if flag and _Side.stream is None and torch.cuda.is_available():
        _Side.stream = torch.cuda.Stream()


def side_join():
    """main stream waits for every weight-gradient kernel issued so far; call after backward()."""
    if _Side.stream is not None and (_Side.pending or _Side.pending_bytes):
        torch.cuda.current_stream().wait_stream(_Side.stream)
    _Side.pending.clear()
    _Side.pending_bytes = 0


def _on_side(weight, fn, keep):
    """weight gradient `fn()` of leaf parameter `weight`: on the main stream it is returned to autograd as usual; with
    the side stream enabled it is computed AND accumulated into weight.grad there (autograd's own accumulation kernels
    would run on the main stream and race with the side stream - e.g. the spatial_reduction convs that are applied
    twice), and autograd gets None for it."""
    if not _Side.enabled or not (weight.is_leaf and weight.requires_grad):
        return fn()
    cur = torch.cuda.current_stream()
    side = _Side.stream
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        dw = fn()
        red = _Side.reducer
        if red is not None and red.side_grad(weight, dw):
            pass                  # written into its bucket slice; the bucket's all-reduce starts when it is complete
        elif weight.grad is None:
            weight.grad = dw
        else:
            weight.grad.add_(dw)
    for t in keep:
        if t is not None:
            _Side.pending.append(t)
            _Side.pending_bytes += t.numel() * t.element_size()
    _Side.pending_bytes += 1
    if _Side.pending_bytes > _Side.limit:
        side_join()
    return None


def _f32(p):
    if p is None:
        return None
    d = p.detach()
    return d if (d.dtype == torch.float32 and d.is_contiguous()) else d.float().contiguous()


def _colsum(g):
    """per-channel sum over pixels of a dense NHWC bf16 tensor -> fp32 [C]"""
    C = g.shape[-1]
    npix = g.numel() // C
    nb = L.lib().dp_chan_reduce_blocks()
    part = torch.empty(nb, 2, C, dtype=torch.float32, device=g.device)
    L.check(L.lib().dp_chan_reduce(0, L.ptr(g), C, None, 0, None, 0, None, npix, C, L.ptr(part), L.stream()))
    out = torch.empty(C, dtype=torch.float32, device=g.device)
    L.check(L.lib().dp_sum_partials(L.ptr(part), nb, 1, C, L.ptr(out), 0, L.stream()))
    return out


def _mask_grad(g_raw, g_relu, y):
    """g_raw + g_relu * (y > 0) (either may be None)."""
    if g_relu is None:
        return _dense(g_raw)
    g_raw, g_relu = _dense(g_raw), _dense(g_relu)
    out = torch.empty(y.shape, dtype=BF16, device=y.device)
    yd = y if y.is_contiguous() else y.contiguous()
    L.check(L.lib().dp_add_relu_bwd(L.ptr(g_raw), L.ptr(g_relu), L.ptr(yd), L.ptr(out), out.numel(), L.stream()))
    return out


# --------------------------------------------------------------------------------------------------
# layout boundaries
# --------------------------------------------------------------------------------------------------
class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, C, H, W = x.shape
        ctx.shape = x.shape
        xs = x.detach()
        if xs.dtype != torch.float32 or not xs.is_contiguous():
            xs = xs.float().contiguous()
        out = _nhwc(B, H, W, C, x.device)
        L.check(L.lib().dp_nchw_f32_to_nhwc_bf16(L.ptr(xs), B, C, H, W, L.ptr(out), C, L.stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        g = _dense(g)
        out = torch.empty(B, C, H, W, dtype=torch.float32, device=g.device)
        L.check(L.lib().dp_nhwc_bf16_to_nchw_f32(L.ptr(g), C, B, C, H, W, L.ptr(out), L.stream()))
        return out


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, H, W, C = x.shape
        ld = _ld(x)
        out = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device)
        L.check(L.lib().dp_nhwc_bf16_to_nchw_f32(L.ptr(x), ld, B, C, H, W, L.ptr(out), L.stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = g.shape
        gs = g if (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous()
        out = _nhwc(B, H, W, C, g.device)
        L.check(L.lib().dp_nchw_f32_to_nhwc_bf16(L.ptr(gs), B, C, H, W, L.ptr(out), C, L.stream()))
        return out


@_internal
def to_nhwc(x):
    """(B,C,H,W) fp32 -> (B,H,W,C) bf16"""
    return _ToNHWC.apply(x)


def to_nchw(x):
    """(B,H,W,C) bf16 -> (B,C,H,W) fp32"""
    return _ToNCHW.apply(x)


@_internal
def tokens_to_nhwc(tok, ph, pw):
    """(B, ph*pw, C) fp32 tokens of the frozen ViT -> (B, ph, pw, C) bf16 (dpt_depth.py:122 without the permute)."""
    B, N, C = tok.shape
    assert N == ph * pw
    src = tok.detach()
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    out = _nhwc(B, ph, pw, C, tok.device)
    L.check(L.lib().dp_cast_f32_to_bf16(L.ptr(src), L.ptr(out), src.numel(), L.stream()))
    return out


# --------------------------------------------------------------------------------------------------
# tensor-core convolution (3x3 s1 p1 / 1x1)
# --------------------------------------------------------------------------------------------------
def _conv_tc_launch(x, wp, Cout, KS, bias, res, res2, relu, out, out2, relu2, stats, fuse=None):
    """the one place a tcgen05 forward / data-gradient convolution is launched (bench.py times this call);
    fuse: an _lib.ConvFuse (BatchNorm prologue / backward-mask epilogue) or None"""
    import ctypes
    B, H, W, Cin = x.shape
    L.check(L.lib().dp_conv2d_tc_fused(
        L.ptr(x), _ld(x), B, H, W, Cin, L.ptr(wp), wp.shape[2], Cout, KS, L.ptr(bias),
        L.ptr(res), _ld(res) if res is not None else 0, L.ptr(res2), _ld(res2) if res2 is not None else 0,
        int(relu), L.ptr(out), _ld(out) if out is not None else 0, L.ptr(out2), _ld(out2) if out2 is not None else 0,
        int(relu2), L.ptr(stats), ctypes.byref(fuse) if fuse is not None else None, L.stream()))


def _wgrad_tc(x, g, Cin, Cout, KS, pre=None):
    """the one place a tcgen05 weight gradient is launched (bench.py times this call); pre = (scale_shift, act): x is
    the pre-BatchNorm tensor and the activated operand is produced in the kernel's operand path"""
    B, H, W, _ = x.shape
    lib = L.lib()
    nb = lib.dp_conv2d_wgrad_tc_workspace(B, H, W, Cin, Cout, KS)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    dw = torch.empty(Cout, Cin, KS, KS, dtype=torch.float32, device=x.device)
    L.check(lib.dp_conv2d_wgrad_tc_fused(L.ptr(x), _ld(x), L.ptr(g), _ld(g), B, H, W, Cin, Cout, KS, L.ptr(dw), 0, L.ptr(ws),
                                         nb, L.ptr(pre[0]) if pre is not None else None,
                                         int(pre[1]) if pre is not None else 0, L.stream()))
    return dw


class _ConvTC(torch.autograd.Function):
    """y = conv(x, w) + b (+ res) (+ res2);  outputs: (relu ? relu(y) : y [, relu(y) copy when dual] [, BN partials])."""

    @staticmethod
    def forward(ctx, x, weight, bias, res, res2, relu, dual, want_stats):
        Cout, Cin, KS, _ = weight.shape
        B, H, W, _ = x.shape
        wp = PACKS.get(weight, 0, 0)
        out = _nhwc(B, H, W, Cout, x.device)
        out2 = _nhwc(B, H, W, Cout, x.device) if dual else None
        stats = None
        if want_stats:
            g = L.lib().dp_conv2d_tc_grid(B, H, W, Cin, Cout, KS)
            stats = torch.empty(g, 2, Cout, dtype=torch.float32, device=x.device)
        _conv_tc_launch(x, wp, Cout, KS, _f32(bias), res, res2, relu, out, out2, True, stats)
        ctx.relu, ctx.dual, ctx.KS = relu, dual, KS
        ctx.save_for_backward(x, weight, out if relu else None, out2 if dual else None)
        ctx.has = (bias is not None, res is not None, res2 is not None)
        ctx.bias_ref = weakref.ref(bias) if (bias is not None and bias.is_leaf and bias.requires_grad) else None
        ctx.mark_non_differentiable(*([stats] if stats is not None else []))
        return out, out2, stats

    @staticmethod
    def backward(ctx, gy, gy2, _gs):
        x, weight, y_relu, y2 = ctx.saved_tensors
        Cout, Cin, KS, _ = weight.shape
        if ctx.relu:
            g = _mask_grad(None, gy, y_relu)            # single ReLU'd output
            if gy2 is not None:
                raise RuntimeError("dual output is only defined for a raw primary output")
        elif ctx.dual and gy2 is not None:
            g = _mask_grad(gy, gy2, y2) if gy is not None else _mask_grad(None, gy2, y2)
        else:
            g = _dense(gy)
        need = ctx.needs_input_grad
        dx = dw = db = None
        if need[0]:
            wd = PACKS.get(weight, 1, 1)
            B, H, W, _ = x.shape
            dx = _nhwc(B, H, W, Cin, x.device)
            _conv_tc_launch(g, wd, Cin, KS, None, None, None, False, dx, None, False, None)
        if need[1]:
            dw = _on_side(weight, lambda: _wgrad_tc(x, g, Cin, Cout, KS).to(weight.dtype), (x, g))
        if ctx.has[0] and need[2]:
            bias = ctx.bias_ref() if ctx.bias_ref is not None else None
            db = _on_side(bias, lambda: _colsum(g), (g,)) if bias is not None else _colsum(g)
        dres = g if (ctx.has[1] and need[3]) else None
        dres2 = g if (ctx.has[2] and need[4]) else None
        return dx, dw, db, dres, dres2, None, None, None


@_internal
def conv_tc(x, weight, bias=None, res=None, res2=None, relu=False, dual=False, stats=False):
    """3x3/s1/p1 or 1x1 conv on tcgen05.  Returns y, or (y, relu(y)) when dual, with BN partials appended when stats."""
    y, y2, st = _ConvTC.apply(x, weight, bias, res, res2, relu, dual, stats)
    r = (y,)
    if dual:
        r = r + (y2,)
    if stats:
        r = r + (st,)
    return r if len(r) > 1 else y


# --------------------------------------------------------------------------------------------------
# strided / transposed convolutions (CUDA-core gather form)
# --------------------------------------------------------------------------------------------------
def _gather(inp, wp, bias, Ho, Wo, Co, KH, KW, stride, pad, transposed, relu=False):
    B, Hi, Wi, Ci = inp.shape
    out = _nhwc(B, Ho, Wo, Co, inp.device)
    L.check(L.lib().dp_conv_gather(L.ptr(inp), _ld(inp), B, Hi, Wi, Ci, L.ptr(wp), L.ptr(bias), L.ptr(out), Co, Ho, Wo,
                                   Co, KH, KW, stride, pad, int(transposed), int(relu), L.stream()))
    return out


def _tc_down2(inp, wp, bias, Ho, Wo, Co, K, pad, stats=False):
    """stride-2 conv rule on tcgen05 (parity planes)"""
    B, Hi, Wi, Ci = inp.shape
    out = _nhwc(B, Ho, Wo, Co, inp.device)
    st = None
    if stats:
        g = L.lib().dp_conv2d_tc_down2_grid(B, Ho, Wo, Ci, Co, K, pad)
        st = torch.empty(g, 2, Co, dtype=torch.float32, device=inp.device)
    L.check(L.lib().dp_conv2d_tc_down2(L.ptr(inp), _ld(inp), B, Hi, Wi, Ci, L.ptr(wp), wp.shape[2], Co, K, pad,
                                       L.ptr(bias), 0, L.ptr(out), Co, Ho, Wo, L.ptr(st), L.stream()))
    return out, st


def _tc_up2(inp, wp, bias, Ho, Wo, Co, K, pad):
    """transposed stride-2 rule on tcgen05 (four output phases)"""
    B, Hi, Wi, Ci = inp.shape
    out = _nhwc(B, Ho, Wo, Co, inp.device)
    L.check(L.lib().dp_conv2d_tc_up2(L.ptr(inp), _ld(inp), B, Hi, Wi, Ci, L.ptr(wp), wp.shape[2], Co, K, pad,
                                     L.ptr(bias), 0, L.ptr(out), Co, Ho, Wo, L.stream()))
    return out


def _wgrad_tc_s2(P, T, K, pad):
    """grad[cp][ct][ky][kx] = sum_p P[p][cp] * T[2p - pad + k][ct] on tcgen05"""
    B, Hp, Wp, Cp = P.shape
    _, Ht, Wt, Ct = T.shape
    lib = L.lib()
    nb = lib.dp_conv2d_wgrad_tc_s2_workspace(B, Hp, Wp, Cp, Ct, K, pad)
    ws = torch.empty(nb, dtype=torch.uint8, device=P.device)
    out = torch.empty(Cp, Ct, K, K, dtype=torch.float32, device=P.device)
    L.check(lib.dp_conv2d_wgrad_tc_s2(L.ptr(P), _ld(P), Hp, Wp, Cp, L.ptr(T), _ld(T), Ht, Wt, Ct, B, K, pad, L.ptr(out), 0,
                                      L.ptr(ws), nb, L.stream()))
    return out


def _tc_stride2_ok(K, stride, pad, Ci, Co):
    return stride == 2 and pad == 1 and K in (3, 4) and Ci % 8 == 0 and Co % 8 == 0


def _wgrad_direct(P, T, KH, KW, stride, pad, perm):
    B, Hp, Wp, Cp = P.shape
    _, Ht, Wt, Ct = T.shape
    lib = L.lib()
    nb = lib.dp_conv_wgrad_direct_workspace(B, Hp, Wp, Cp, Ct, KH, KW)
    ws = torch.empty(nb, dtype=torch.uint8, device=P.device)
    out = torch.empty(Cp, Ct, KH, KW, dtype=torch.float32, device=P.device)
    L.check(lib.dp_conv_wgrad_direct(L.ptr(P), _ld(P), Hp, Wp, Cp, L.ptr(T), _ld(T), Ht, Wt, Ct, B, KH, KW, stride, pad,
                                     perm, L.ptr(out), 0, L.ptr(ws), nb, L.stream()))
    return out


class _ConvStrided(torch.autograd.Function):
    """nn.Conv2d with stride > 1 (weight [O][I][KH][KW])."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad):
        O, I, KH, KW = weight.shape
        B, Hi, Wi, _ = x.shape
        Ho = (Hi + 2 * pad - KH) // stride + 1
        Wo = (Wi + 2 * pad - KW) // stride + 1
        if KH == KW and _tc_stride2_ok(KH, stride, pad, I, O):
            out, _ = _tc_down2(x, PACKS.get(weight, 0, 0), _f32(bias), Ho, Wo, O, KH, pad)
        else:
            out = _gather(x, PACKS.get(weight, 0, 0), _f32(bias), Ho, Wo, O, KH, KW, stride, pad, False)
        ctx.save_for_backward(x, weight)
        ctx.geom = (stride, pad, bias is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        stride, pad, has_bias = ctx.geom
        O, I, KH, KW = weight.shape
        B, Hi, Wi, _ = x.shape
        g = _dense(g)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if KH == KW and _tc_stride2_ok(KH, stride, pad, O, I):
                dx = _tc_up2(g, PACKS.get(weight, 1, 0), None, Hi, Wi, I, KH, pad)
            else:
                dx = _gather(g, PACKS.get(weight, 1, 0), None, Hi, Wi, I, KH, KW, stride, pad, True)
        if ctx.needs_input_grad[1]:
            if KH == KW and _tc_stride2_ok(KH, stride, pad, I, O):
                dw = _on_side(weight, lambda: _wgrad_tc_s2(g, x, KH, pad).to(weight.dtype), (x, g))
            else:
                dw = _on_side(weight, lambda: _wgrad_direct(g, x, KH, KW, stride, pad, 1).to(weight.dtype), (x, g))  # [O][I][KH][KW]
        if has_bias and ctx.needs_input_grad[2]:
            db = _colsum(g)
        return dx, dw, db, None, None


class _ConvTransposed(torch.autograd.Function):
    """nn.ConvTranspose2d (weight [I][O][KH][KW], output_padding 0)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad):
        I, O, KH, KW = weight.shape
        B, Hi, Wi, _ = x.shape
        Ho = (Hi - 1) * stride - 2 * pad + KH
        Wo = (Wi - 1) * stride - 2 * pad + KW
        if KH == KW and _tc_stride2_ok(KH, stride, pad, I, O):
            out = _tc_up2(x, PACKS.get(weight, 1, 0), _f32(bias), Ho, Wo, O, KH, pad)
        else:
            out = _gather(x, PACKS.get(weight, 1, 0), _f32(bias), Ho, Wo, O, KH, KW, stride, pad, True)
        ctx.save_for_backward(x, weight)
        ctx.geom = (stride, pad, bias is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        stride, pad, has_bias = ctx.geom
        I, O, KH, KW = weight.shape
        B, Hi, Wi, _ = x.shape
        g = _dense(g)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if KH == KW and _tc_stride2_ok(KH, stride, pad, O, I):
                dx, _ = _tc_down2(g, PACKS.get(weight, 0, 0), None, Hi, Wi, I, KH, pad)
            else:
                dx = _gather(g, PACKS.get(weight, 0, 0), None, Hi, Wi, I, KH, KW, stride, pad, False)
        if ctx.needs_input_grad[1]:
            if KH == KW and _tc_stride2_ok(KH, stride, pad, I, O):
                dw = _on_side(weight, lambda: _wgrad_tc_s2(x, g, KH, pad).to(weight.dtype), (x, g))
            else:
                dw = _on_side(weight, lambda: _wgrad_direct(x, g, KH, KW, stride, pad, 1).to(weight.dtype), (x, g))  # [I][O][KH][KW]
        if has_bias and ctx.needs_input_grad[2]:
            db = _colsum(g)
        return dx, dw, db, None, None


@_internal
def conv_strided(x, weight, bias, stride, pad):
    return _ConvStrided.apply(x, weight, bias, stride, pad)


@_internal
def conv_transposed(x, weight, bias, stride, pad):
    I, O, KH, KW = weight.shape
    if KH == KW == stride and pad == 0 and I % 8 == 0 and O % 8 == 0:
        # non-overlapping taps (Dinov2Head.resize_layers[0..1], dpt_depth.py:49-62): every output pixel (k*y+a, k*x+b)
        # sees exactly one tap, so the layer is a 1x1 convolution to k*k*O channels on tcgen05 followed by a pixel
        # shuffle; autograd derives the data / weight gradients as the 1x1 kernels' and routes them back through
        # the weight re-view.
        k = KH
        B, H, W, _ = x.shape
        w1 = weight.permute(2, 3, 1, 0).reshape(k * k * O, I, 1, 1)
        b1 = bias.repeat(k * k) if bias is not None else None
        t = conv_tc(x, w1, b1)                                             # (B, H, W, k*k*O)
        return t.view(B, H, W, k, k, O).permute(0, 1, 3, 2, 4, 5).reshape(B, H * k, W * k, O)
    return _ConvTransposed.apply(x, weight, bias, stride, pad)


# --------------------------------------------------------------------------------------------------
# depthwise convolution and the 3-channel stem of the EfficientNet-Lite3 trunk
# --------------------------------------------------------------------------------------------------
def _dw_pack(weight, flip=False):
    """[C][1][K][K] fp32 -> tap-major fp32 [K*K][C] (optionally spatially flipped: the stride-1 data gradient)."""
    C, _, K, _ = weight.shape
    w = weight.detach().float()
    if flip:
        w = w.flip(2, 3)
    return w.reshape(C, K * K).t().contiguous()


def _dw_launch(x, wp, K, stride, pad_t, pad_l, Ho, Wo, want_stats):
    B, Hi, Wi, C = x.shape
    lib = L.lib()
    out = _nhwc(B, Ho, Wo, C, x.device)
    st = None
    if want_stats:
        st = torch.empty(lib.dp_dwconv_fwd_blocks_s(B, Ho, Wo, C, K, stride), 2, C, dtype=torch.float32, device=x.device)
    L.check(lib.dp_dwconv_fwd(L.ptr(x), _ld(x), B, Hi, Wi, C, L.ptr(wp), K, stride, pad_t, pad_l, L.ptr(out), C, Ho, Wo,
                              L.ptr(st), L.stream()))
    return out, st


class _DwConv(torch.autograd.Function):
    """nn.Conv2d(C, C, K, stride, groups=C, bias=False) on NHWC bf16; optional BN partial statistics of the output."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad_t, pad_l, Ho, Wo, want_stats):
        K = weight.shape[2]
        out, st = _dw_launch(x, _dw_pack(weight), K, stride, pad_t, pad_l, Ho, Wo, want_stats)
        ctx.save_for_backward(x, weight)
        ctx.geom = (stride, pad_t, pad_l, Ho, Wo)
        if st is not None:
            ctx.mark_non_differentiable(st)
        return out, st

    @staticmethod
    def backward(ctx, g, _gs):
        x, weight = ctx.saved_tensors
        stride, pad_t, pad_l, Ho, Wo = ctx.geom
        B, Hi, Wi, C = x.shape
        K = weight.shape[2]
        lib = L.lib()
        g = _dense(g)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if stride == 1:
                dx, _ = _dw_launch(g, _dw_pack(weight, flip=True), K, 1, K - 1 - pad_t, K - 1 - pad_l, Hi, Wi, False)
            else:
                dx = _nhwc(B, Hi, Wi, C, x.device)
                L.check(lib.dp_dwconv_dgrad_s2(L.ptr(g), C, B, Ho, Wo, C, L.ptr(_dw_pack(weight)), K, pad_t, pad_l,
                                               L.ptr(dx), C, Hi, Wi, L.stream()))
        if ctx.needs_input_grad[1]:
            def wgrad():
                nb = lib.dp_dwconv_wgrad_workspace(B, Ho, Wo, C, K)
                ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
                out = torch.empty(C, 1, K, K, dtype=torch.float32, device=x.device)
                L.check(lib.dp_dwconv_wgrad(L.ptr(x), _ld(x), B, Hi, Wi, C, L.ptr(g), C, Ho, Wo, K, stride, pad_t, pad_l,
                                            L.ptr(out), 0, L.ptr(ws), nb, L.stream()))
                return out.to(weight.dtype)
            dw = _on_side(weight, wgrad, (x, g))
        return dx, dw, None, None, None, None, None, None


@_internal
def dwconv(x, weight, stride, pad_t, pad_l, Ho, Wo, stats=False):
    """depthwise conv; returns out or (out, BN partials)"""
    out, st = _DwConv.apply(x, weight, stride, pad_t, pad_l, Ho, Wo, stats)
    return (out, st) if stats else out


class _StemConv(torch.autograd.Function):
    """3x3 / stride 2 / pad 1 convolution of the fp32 NCHW image (3 channels, zero padded to 8 for the tensor cores)."""

    @staticmethod
    def forward(ctx, x, weight, want_stats):
        B, Ci, H, W = x.shape
        O, _, K, _ = weight.shape
        assert K == 3 and Ci <= 8
        xs = x.detach()
        if xs.dtype != torch.float32 or not xs.is_contiguous():
            xs = xs.float().contiguous()
        x8 = torch.empty(B, H, W, 8, dtype=BF16, device=x.device)     # pad channels are zeroed by the converter
        L.check(L.lib().dp_nchw_f32_to_nhwc_bf16(L.ptr(xs), B, Ci, H, W, L.ptr(x8), 8, L.stream()))
        wp = torch.zeros(K * K, O, 8, dtype=BF16, device=x.device)
        wp[:, :, :Ci] = weight.detach().permute(2, 3, 0, 1).reshape(K * K, O, Ci).to(BF16)
        Ho, Wo = (H + 2 - K) // 2 + 1, (W + 2 - K) // 2 + 1
        out, st = _tc_down2(x8, wp, None, Ho, Wo, O, K, 1, stats=want_stats)
        ctx.save_for_backward(x8, weight)
        if st is not None:
            ctx.mark_non_differentiable(st)
        return out, st

    @staticmethod
    def backward(ctx, g, _gs):
        x8, weight = ctx.saved_tensors
        Ci = weight.shape[1]
        dw = _wgrad_tc_s2(_dense(g), x8, 3, 1)[:, :Ci].contiguous().to(weight.dtype)
        return None, dw, None


@_internal
def stem_conv(x, weight, stats=False):
    out, st = _StemConv.apply(x, weight, stats)
    return (out, st) if stats else out


# --------------------------------------------------------------------------------------------------
# C -> 1 head conv (+bias, +ReLU) producing the fp32 (B,H,W) depth map
# --------------------------------------------------------------------------------------------------
class _HeadConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        _, C, KS, _ = weight.shape
        B, H, W, _ = x.shape
        out = torch.empty(B, H, W, dtype=torch.float32, device=x.device)
        L.check(L.lib().dp_head_conv_fwd(L.ptr(x), _ld(x), B, H, W, C, KS, L.ptr(_f32(weight)), L.ptr(_f32(bias)),
                                         int(relu), L.ptr(out), L.stream()))
        ctx.relu = relu
        ctx.save_for_backward(x, weight, out)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, out = ctx.saved_tensors
        _, C, KS, _ = weight.shape
        B, H, W, _ = x.shape
        g = g if (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous()
        lib = L.lib()
        nb = lib.dp_head_conv_bwd_workspace(C, KS)
        ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
        dx = _nhwc(B, H, W, C, x.device) if ctx.needs_input_grad[0] else None
        dw = torch.empty(1, C, KS, KS, dtype=torch.float32, device=x.device)
        db = torch.empty(1, dtype=torch.float32, device=x.device)
        L.check(lib.dp_head_conv_bwd(L.ptr(g), L.ptr(out), int(ctx.relu), L.ptr(x), _ld(x), B, H, W, C, KS,
                                     L.ptr(_f32(weight)), L.ptr(dx), C, L.ptr(dw), L.ptr(db), 0, L.ptr(ws), nb, L.stream()))
        return dx, dw.to(weight.dtype), (db if ctx.has_bias else None), None


def head_conv(x, weight, bias, relu):
    return _HeadConv.apply(x, weight, bias, relu)


# --------------------------------------------------------------------------------------------------
# bilinear resize
# --------------------------------------------------------------------------------------------------
class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Ho, Wo, align):
        B, Hi, Wi, C = x.shape
        out = _nhwc(B, Ho, Wo, C, x.device)
        L.check(L.lib().dp_resize_bilinear_nhwc(L.ptr(x), _ld(x), B, Hi, Wi, C, L.ptr(out), C, Ho, Wo, int(align),
                                                L.stream()))
        ctx.geom = (B, Hi, Wi, C, Ho, Wo, align)
        return out

    @staticmethod
    def backward(ctx, g):
        B, Hi, Wi, C, Ho, Wo, align = ctx.geom
        g = _dense(g)
        gin = _nhwc(B, Hi, Wi, C, g.device)
        L.check(L.lib().dp_resize_bilinear_nhwc_bwd(L.ptr(g), C, B, Hi, Wi, C, L.ptr(gin), C, Ho, Wo, int(align),
                                                    L.stream()))
        return gin, None, None, None


@_internal
def resize(x, size, align_corners):
    Ho, Wo = int(size[0]), int(size[1])
    if (Ho, Wo) == tuple(x.shape[1:3]):
        return x
    return _Resize.apply(x, Ho, Wo, bool(align_corners))


def resize_planes_f32(x, size, align_corners):
    """(B,C,H,W) fp32 -> (B,C,Ho,Wo) fp32, no gradient (RGB -> frozen DINOv2 input; prediction resize in eval)."""
    B, C, Hi, Wi = x.shape
    Ho, Wo = int(size[0]), int(size[1])
    src = x.detach()
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    out = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=x.device)
    L.check(L.lib().dp_resize_bilinear_planes_f32(L.ptr(src), B * C, Hi, Wi, L.ptr(out), Ho, Wo, int(align_corners),
                                                  L.stream()))
    return out


# --------------------------------------------------------------------------------------------------
# BatchNorm2d (+ residual / second normalised branch) (+ ReLU)
# --------------------------------------------------------------------------------------------------
def _bn_coeffs(bn, c, stats, training):
    """returns (scale_shift [2][C], save [2][C] = mean, invstd); updates running stats in train mode."""
    C = c.shape[-1]
    dev = c.device
    lib = L.lib()
    ss = torch.empty(2, C, dtype=torch.float32, device=dev)
    save = torch.empty(2, C, dtype=torch.float32, device=dev)
    gamma, beta = _f32(bn.weight), _f32(bn.bias)
    use_batch = training or bn.running_mean is None
    if use_batch:
        if stats is None:
            npix = c.numel() // C
            nb = lib.dp_chan_reduce_blocks()
            stats = torch.empty(nb, 2, C, dtype=torch.float32, device=dev)
            cd = c if c.is_contiguous() else c.contiguous()
            L.check(lib.dp_chan_reduce(1, L.ptr(cd), C, None, 0, None, 0, None, npix, C, L.ptr(stats), L.stream()))
        count = float(c.numel() // C)
        track = training and bn.track_running_stats and bn.running_mean is not None
        mom = bn.momentum if bn.momentum is not None else 0.1
        L.check(lib.dp_bn_finalize(L.ptr(stats), stats.shape[0], C, count, L.ptr(gamma), L.ptr(beta), bn.eps, mom,
                                   L.ptr(bn.running_mean) if track else None, L.ptr(bn.running_var) if track else None,
                                   L.ptr(bn.num_batches_tracked) if track else None, L.ptr(ss), L.ptr(save), L.stream()))
    else:
        L.check(lib.dp_bn_eval_coeffs(L.ptr(gamma), L.ptr(beta), L.ptr(bn.running_mean), L.ptr(bn.running_var), bn.eps,
                                      C, L.ptr(ss), L.ptr(save), L.stream()))
    return ss, save, use_batch


class _BNAct(torch.autograd.Function):
    """y = act( bn(c) [+ bn2(c2) | + res] ).  Parameters enter as tensors so autograd routes their grads."""

    @staticmethod
    def forward(ctx, c, gamma, beta, ss, save, c2, gamma2, beta2, ss2, save2, res, relu, train):
        B, H, W, C = c.shape
        npix = B * H * W
        y = _nhwc(B, H, W, C, c.device)
        relu = int(relu)                      # 0 none, 1 ReLU, 2 ReLU6
        L.check(L.lib().dp_bn_apply(L.ptr(c), _ld(c), L.ptr(ss), L.ptr(c2), _ld(c2) if c2 is not None else 0,
                                    L.ptr(ss2), L.ptr(res), _ld(res) if res is not None else 0, npix, C, relu,
                                    L.ptr(y), C, L.stream()))
        # BN (+ residual) + activation: the backward recomputes the activation mask in fp32 from c, the scale/shift and
        # the residual; only with a second normalised branch is the stored output used as the mask
        recompute = bool(relu) and c2 is None
        ctx.save_for_backward(c, gamma, save, c2, gamma2, save2,
                              (res if recompute else y) if relu else None, ss if recompute else None)
        ctx.cfg = (relu, train, res is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        c, gamma, save, c2, gamma2, save2, y, mss = ctx.saved_tensors
        relu, train, has_res = ctx.cfg
        B, H, W, C = c.shape
        npix = B * H * W
        lib = L.lib()
        gy = _dense(gy)
        dev = c.device
        nb = lib.dp_chan_reduce_blocks()

        def branch(cx, gm, sv, want_gmask):
            part = torch.empty(nb, 2, C, dtype=torch.float32, device=dev)
            L.check(lib.dp_chan_reduce(3 if relu == 2 else 2, L.ptr(cx), _ld(cx), L.ptr(gy), C, L.ptr(y), C, L.ptr(mss),
                                       npix, C, L.ptr(part), L.stream()))
            red = torch.empty(2, C, dtype=torch.float32, device=dev)
            L.check(lib.dp_sum_partials(L.ptr(part), nb, 2, C, L.ptr(red), 0, L.stream()))
            dx = _nhwc(B, H, W, C, dev)
            gmask = _nhwc(B, H, W, C, dev) if want_gmask else None
            dg = torch.empty(C, dtype=torch.float32, device=dev)
            dbt = torch.empty(C, dtype=torch.float32, device=dev)
            L.check(lib.dp_bn_bwd_apply(L.ptr(gy), C, L.ptr(y), C, L.ptr(mss), L.ptr(cx), _ld(cx), L.ptr(red), L.ptr(sv),
                                        L.ptr(_f32(gm)), float(npix), int(train) | (2 if relu == 2 else 0), npix, C,
                                        L.ptr(dx), C, L.ptr(gmask), C,
                                        L.ptr(dg), L.ptr(dbt), 0, L.stream()))
            return dx, dg, dbt, gmask

        dc, dg, db, gmask = branch(c, gamma, save, has_res)
        dc2 = dg2 = db2 = None
        if c2 is not None:
            dc2, dg2, db2, _ = branch(c2, gamma2, save2, False)
        return (dc, dg.to(gamma.dtype), db.to(gamma.dtype), None, None, dc2,
                dg2.to(gamma2.dtype) if dg2 is not None else None, db2.to(gamma2.dtype) if db2 is not None else None,
                None, None, gmask, None, None)


@_internal
def bn_act(bn, c, stats=None, relu=True, res=None, bn2=None, c2=None, stats2=None):
    """train/eval nn.BatchNorm2d on NHWC bf16 `c` (+ residual, or + a second normalised branch), optional ReLU
    (relu=True / 1) or ReLU6 (relu=2)."""
    training = bn.training
    ss, save, used_batch = _bn_coeffs(bn, c, stats, training)
    ss2 = save2 = g2 = b2 = None
    if bn2 is not None:
        ss2, save2, _ = _bn_coeffs(bn2, c2, stats2, bn2.training)
        g2, b2 = bn2.weight, bn2.bias
    return _BNAct.apply(c, bn.weight, bn.bias, ss, save, c2, g2, b2, ss2, save2, res, relu, used_batch)


# --------------------------------------------------------------------------------------------------
# ResidualBlock as ONE autograd node: BatchNorm folded into the neighbouring convolutions
# --------------------------------------------------------------------------------------------------
def _conv_raw(x, wp, Cout, KS, res=None, stats=False, pre=None, mask=None):
    """plain launch of the tcgen05 convolution (no autograd): returns (out, stats partials | None).
    pre = (scale_shift [2][Cin], act): BatchNorm + activation of the input fused into the operand path;
    mask = (c, scale_shift [2][Cout], act): BatchNorm-backward epilogue (masked gradient + its two batch sums)."""
    B, H, W, Cin = x.shape
    lib = L.lib()
    out = _nhwc(B, H, W, Cout, x.device)
    st = None
    if stats or mask is not None:
        st = torch.empty(lib.dp_conv2d_tc_grid(B, H, W, Cin, Cout, KS), 2, Cout, dtype=torch.float32, device=x.device)
    fuse = None
    if pre is not None or mask is not None:
        fuse = L.ConvFuse(L.ptr(pre[0]) if pre is not None else None, int(pre[1]) if pre is not None else 0,
                          L.ptr(mask[0]) if mask is not None else None, _ld(mask[0]) if mask is not None else 0,
                          L.ptr(mask[1]) if mask is not None else None, int(mask[2]) if mask is not None else 0)
    _conv_tc_launch(x, wp, Cout, KS, None, res, None, False, out, None, False, st, fuse)
    return out, st


def _wgrad_raw(x, g, Cin, Cout, KS, pre=None):
    return _wgrad_tc(x, g, Cin, Cout, KS, pre)


def _bn_reduce(cx, gy, mask, mss, act):
    """red[2][C] = (sum g, sum g*c) with g = gy * [activation passed]; the activation is recomputed from c (mss, mask =
    residual added before it) or read from the stored output (mask = y, mss None)."""
    lib = L.lib()
    C = cx.shape[-1]
    npix = cx.numel() // C
    nb = lib.dp_chan_reduce_blocks()
    part = torch.empty(nb, 2, C, dtype=torch.float32, device=cx.device)
    L.check(lib.dp_chan_reduce(3 if act == 2 else 2, L.ptr(cx), _ld(cx), L.ptr(gy), _ld(gy), L.ptr(mask),
                               _ld(mask) if mask is not None else 0, L.ptr(mss), npix, C, L.ptr(part), L.stream()))
    return _fold_partials(part, C)


def _fold_partials(part, C):
    red = torch.empty(2, C, dtype=torch.float32, device=part.device)
    L.check(L.lib().dp_sum_partials(L.ptr(part), part.shape[0], 2, C, L.ptr(red), 0, L.stream()))
    return red


def _bn_bwd(gy, mask, mss, cx, red, save, gamma, train, act, want_gmask=False, want_dx=True):
    """BatchNorm backward apply: returns (dc, dgamma, dbeta, gmask)"""
    lib = L.lib()
    B, H, W, C = cx.shape
    npix = B * H * W
    dev = cx.device
    dx = _nhwc(B, H, W, C, dev) if want_dx else None
    gmask = _nhwc(B, H, W, C, dev) if want_gmask else None
    dg = torch.empty(C, dtype=torch.float32, device=dev)
    db = torch.empty(C, dtype=torch.float32, device=dev)
    L.check(lib.dp_bn_bwd_apply(L.ptr(gy), _ld(gy), L.ptr(mask), _ld(mask) if mask is not None else 0, L.ptr(mss),
                                L.ptr(cx), _ld(cx), L.ptr(red), L.ptr(save), L.ptr(_f32(gamma)), float(npix),
                                int(train) | (2 if act == 2 else 0), npix, C, L.ptr(dx), C, L.ptr(gmask), C, L.ptr(dg),
                                L.ptr(db), 0, L.stream()))
    return dx, dg, db, gmask


def _bn_apply_raw(c, ss, act, res=None, c2=None, ss2=None):
    B, H, W, C = c.shape
    y = _nhwc(B, H, W, C, c.device)
    L.check(L.lib().dp_bn_apply(L.ptr(c), _ld(c), L.ptr(ss), L.ptr(c2), _ld(c2) if c2 is not None else 0, L.ptr(ss2),
                                L.ptr(res), _ld(res) if res is not None else 0, B * H * W, C, int(act), L.ptr(y), C,
                                L.stream()))
    return y


class Fusion:
    """which BatchNorm fusions the block-level autograd nodes use.  Policy from tools/fused_micro.py on B200 (B = 32,
    448x576, ms per launch; profiles/fused_micro_r2.txt):

      backward (ReLU mask + BatchNorm-backward sums in the data-gradient epilogue): 64->64 0.89 vs 0.52 + 0.45 for the
        separate reduction pass, 32->32 0.45 vs 0.23 + 0.23, 32->64 0.78 vs 0.37 + 0.45: on (break-even to 8 % faster,
        one launch and one full read of the gradient fewer).
      prologue (bn + ReLU inside the consumer conv's operand path): 64->64 forward 0.76 vs 0.52 + 0.41 for conv + bn_apply,
        but the weight gradient then has to re-create the activation in its own operand path (0.99 vs 0.68), and at
        <= 32 channels the four transform warps cannot keep up with the HBM-bound tile rate (32->32 0.53 vs 0.23 + 0.21).
        "auto": used where it wins - forward passes that keep no graph (eval / no_grad) with >= 64 input channels;
        True / False force it (tests compare both paths bit for bit)."""
    prologue = "auto"
    backward = True

    @staticmethod
    def use_prologue(cin, needs_grad):
        if Fusion.prologue == "auto":
            return (not needs_grad) and cin >= 64
        return bool(Fusion.prologue)


class _ResBlock(torch.autograd.Function):
    """relu(bn2(conv2(relu(bn1(conv1(x))))) + shortcut(x))  (reference midas_semantics.py:129-151) as one autograd node.

    forward:  conv1 (+ BN partial sums in its epilogue) -> conv2 whose operand path applies bn1 + ReLU to the landed
              tiles (the activated tensor never reaches HBM; its BN partial sums again come from the epilogue) ->
              one pass for bn2 + shortcut + ReLU.
    backward: bn2 backward (reduce + apply) -> conv2 data gradient whose epilogue masks with bn1's ReLU and produces bn1's
              two batch sums -> bn1 backward apply -> conv1 data gradient with the shortcut's gradient added in its
              epilogue (no autograd accumulation kernel); the weight gradient of conv2 re-creates relu(bn1(c1)) in its
              own operand path."""

    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, ws, gs, bs, bn1, bn2, bns):
        C1, Cin = w1.shape[0], w1.shape[1]
        C2 = w2.shape[0]
        train = bn1.training
        c1, st1 = _conv_raw(x, PACKS.get(w1, 0, 0), C1, 3, stats=train)
        ss1, save1, tr1 = _bn_coeffs(bn1, c1, st1, train)
        a1 = None
        if Fusion.use_prologue(C1, any(ctx.needs_input_grad)):
            c2, st2 = _conv_raw(c1, PACKS.get(w2, 0, 0), C2, 3, stats=bn2.training, pre=(ss1, 1))
        else:
            a1 = _bn_apply_raw(c1, ss1, 1)
            c2, st2 = _conv_raw(a1, PACKS.get(w2, 0, 0), C2, 3, stats=bn2.training)
        ss2, save2, tr2 = _bn_coeffs(bn2, c2, st2, bn2.training)
        cs = sss = saves = None
        trs = False
        if ws is None:
            y = _bn_apply_raw(c2, ss2, 1, res=x)
        else:
            cs, sts = _conv_raw(x, PACKS.get(ws, 0, 0), C2, 1, stats=bns.training)
            sss, saves, trs = _bn_coeffs(bns, cs, sts, bns.training)
            y = _bn_apply_raw(c2, ss2, 1, c2=cs, ss2=sss)
        ctx.save_for_backward(x, w1, g1, w2, g2, ws, gs, c1, c2, cs, y if ws is not None else None, ss1, save1, ss2, save2,
                              sss, saves, a1)
        ctx.trains = (tr1, tr2, trs)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w1, g1, w2, g2, ws, gs, c1, c2, cs, y, ss1, save1, ss2, save2, sss, saves, a1 = ctx.saved_tensors
        tr1, tr2, trs = ctx.trains
        C1, Cin = w1.shape[0], w1.shape[1]
        C2 = w2.shape[0]
        B, H, W, _ = x.shape
        gy = _dense(gy)
        need_x = ctx.needs_input_grad[0]
        # ---- bn2 (+ shortcut bn) ----
        gskip = dcs = dgs = dbs = None
        if ws is None:
            red2 = _bn_reduce(c2, gy, x, ss2, 1)
            dc2, dg2, db2, gskip = _bn_bwd(gy, x, ss2, c2, red2, save2, g2, tr2, 1, want_gmask=need_x)
        else:
            red2 = _bn_reduce(c2, gy, y, None, 1)
            dc2, dg2, db2, _ = _bn_bwd(gy, y, None, c2, red2, save2, g2, tr2, 1)
            reds = _bn_reduce(cs, gy, y, None, 1)
            dcs, dgs, dbs, _ = _bn_bwd(gy, y, None, cs, reds, saves, gs, trs, 1)
        # ---- conv2: data gradient with bn1's ReLU mask and batch sums in the epilogue ----
        wd2 = PACKS.get(w2, 1, 1)
        if Fusion.backward and L.lib().dp_conv2d_tc_caps(B, H, W, C2, C1, 3) & L.CAP_BN_BACKWARD:
            dz1, part = _conv_raw(dc2, wd2, C1, 3, mask=(c1, ss1, 1))
            red1 = _fold_partials(part, C1)
            dc1, dg1, db1, _ = _bn_bwd(dz1, None, None, c1, red1, save1, g1, tr1, 0)
        else:
            da1, _ = _conv_raw(dc2, wd2, C1, 3)
            red1 = _bn_reduce(c1, da1, None, ss1, 1)
            dc1, dg1, db1, _ = _bn_bwd(da1, None, ss1, c1, red1, save1, g1, tr1, 1)
        dw2 = None
        if ctx.needs_input_grad[4]:
            if a1 is None:
                dw2 = _on_side(w2, lambda: _wgrad_raw(c1, dc2, C1, C2, 3, pre=(ss1, 1)).to(w2.dtype), (c1, dc2))
            else:
                dw2 = _on_side(w2, lambda: _wgrad_raw(a1, dc2, C1, C2, 3).to(w2.dtype), (a1, dc2))
        # ---- conv1 (+ shortcut): the skip gradient rides in the data-gradient epilogue ----
        dx = dw1 = dws = None
        if need_x:
            if ws is not None:
                gskip, _ = _conv_raw(dcs, PACKS.get(ws, 1, 1), Cin, 1)
            dx, _ = _conv_raw(dc1, PACKS.get(w1, 1, 1), Cin, 3, res=gskip)
        if ctx.needs_input_grad[1]:
            dw1 = _on_side(w1, lambda: _wgrad_raw(x, dc1, Cin, C1, 3).to(w1.dtype), (x, dc1))
        if ws is not None and ctx.needs_input_grad[7]:
            dws = _on_side(ws, lambda: _wgrad_raw(x, dcs, Cin, C2, 1).to(ws.dtype), (x, dcs))
        cast = lambda t, ref: t.to(ref.dtype) if t is not None else None
        return (dx, dw1, cast(dg1, g1), cast(db1, g1), dw2, cast(dg2, g2), cast(db2, g2), dws,
                cast(dgs, gs) if gs is not None else None, cast(dbs, gs) if gs is not None else None, None, None, None)


@_internal
def res_block(x, conv1, bn1, conv2, bn2, shortcut):
    """fused ResidualBlock (stride 1, bias-free convs); shortcut = None (identity) or (conv1x1, bn)."""
    ws = gs = bs = bns = None
    if shortcut is not None:
        ws, bns = shortcut[0].weight, shortcut[1]
        gs, bs = bns.weight, bns.bias
    return _ResBlock.apply(x, conv1.weight, bn1.weight, bn1.bias, conv2.weight, bn2.weight, bn2.bias, ws, gs, bs,
                           bn1, bn2, bns)


# --------------------------------------------------------------------------------------------------
# LayerNorm + Linear, segmented attention
# --------------------------------------------------------------------------------------------------
class _LNLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, W, bias, eps, out_bf16):
        # x: (..., 32) bf16 NHWC tokens or fp32 [B][N][32]
        dim = x.shape[-1]
        x_f32 = x.dtype == torch.float32
        xs = x if x.is_contiguous() else x.contiguous()
        ntok = xs.numel() // dim
        out = torch.empty(x.shape, dtype=BF16 if out_bf16 else torch.float32, device=x.device)
        L.check(L.lib().dp_ln_linear_fwd(L.ptr(xs), dim, int(x_f32), ntok, dim, L.ptr(_f32(gamma)), L.ptr(_f32(beta)), eps,
                                         L.ptr(_f32(W)), L.ptr(_f32(bias)), L.ptr(out), dim, int(out_bf16), L.stream()))
        ctx.save_for_backward(xs, gamma, beta, W)
        ctx.cfg = (eps, bias is not None, x_f32, out_bf16)
        return out

    @staticmethod
    def backward(ctx, g):
        xs, gamma, beta, W = ctx.saved_tensors
        eps, has_bias, x_f32, out_bf16 = ctx.cfg
        dim = xs.shape[-1]
        ntok = xs.numel() // dim
        lib = L.lib()
        want = BF16 if out_bf16 else torch.float32
        g = g if (g.dtype == want and g.is_contiguous()) else g.to(want).contiguous()
        dev = xs.device
        part = torch.empty(lib.dp_lnl_blocks(), lib.dp_lnl_partial_floats(), dtype=torch.float32, device=dev)
        dx = torch.empty_like(xs) if ctx.needs_input_grad[0] else None
        dW = torch.empty(dim, dim, dtype=torch.float32, device=dev)
        dbias = torch.empty(dim, dtype=torch.float32, device=dev) if has_bias else None
        dg = torch.empty(dim, dtype=torch.float32, device=dev)
        db = torch.empty(dim, dtype=torch.float32, device=dev)
        L.check(lib.dp_ln_linear_bwd(L.ptr(xs), dim, int(x_f32), ntok, dim, L.ptr(_f32(gamma)), L.ptr(_f32(beta)), eps,
                                     L.ptr(_f32(W)), L.ptr(g), dim, int(out_bf16), L.ptr(dx), dim, L.ptr(part), L.ptr(dW),
                                     L.ptr(dbias), L.ptr(dg), L.ptr(db), 0, L.stream()))
        return dx, dg.to(gamma.dtype), db.to(beta.dtype), dW.to(W.dtype), dbias, None, None


@_internal
def ln_linear(x, ln, lin, out_bf16=False):
    return _LNLinear.apply(x, ln.weight, ln.bias, lin.weight, lin.bias, ln.eps, out_bf16)


_SEG_CACHE = {}


def attention_segments(hr, wr, ws, device):
    """The reference's window loop (midas_semantics.py:93-112) reduced to its last-writer form.
    Returns (items int32 [n,4] {q0, nq<=32, k_lo, k_hi}, segs int32 [m,4] {q_lo, q_hi, k_lo, k_hi})."""
    key = (hr, wr, ws, str(device))
    if key in _SEG_CACHE:
        return _SEG_CACHE[key]
    N = hr * wr
    ranges = []
    for h in range((hr + ws - 1) // ws):
        for w in range((wr + ws - 1) // ws):
            lo = h * ws * wr + w * ws
            hi = min(min(h * ws + ws, hr) * wr + min(w * ws + ws, wr), N)
            ranges.append((lo, hi))
    owner = [-1] * N
    for wi, (lo, hi) in enumerate(ranges):
        for i in range(lo, hi):
            owner[i] = wi
    segs = []
    i = 0
    while i < N:
        j = i
        while j < N and owner[j] == owner[i]:
            j += 1
        if owner[i] >= 0:
            lo, hi = ranges[owner[i]]
            segs.append((i, j, lo, hi))
        i = j
    items = []
    for (a, b, lo, hi) in segs:
        for q0 in range(a, b, 32):
            items.append((q0, min(32, b - q0), lo, hi))
    unowned = [i for i in range(N) if owner[i] < 0]
    res = (torch.tensor(items, dtype=torch.int32, device=device), torch.tensor(segs, dtype=torch.int32, device=device),
           unowned)
    _SEG_CACHE[key] = res
    return res


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, items, segs, scale):
        B, N, D = q.shape
        out = torch.zeros(B, N, D, dtype=torch.float32, device=q.device)
        lse = torch.zeros(B, N, 8, dtype=torch.float32, device=q.device)
        L.check(L.lib().dp_attn_fwd(L.ptr(q), L.ptr(k), L.ptr(v), B, N, scale, L.ptr(items), items.shape[0], L.ptr(out),
                                    L.ptr(lse), L.stream()))
        ctx.save_for_backward(q, k, v, out, lse, items, segs)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, g):
        q, k, v, out, lse, items, segs = ctx.saved_tensors
        B, N, D = q.shape
        g = g if (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous()
        dq = torch.zeros_like(q)
        dk = torch.zeros_like(k)
        dv = torch.zeros_like(v)
        delta = torch.zeros(B, N, 8, dtype=torch.float32, device=q.device)
        L.check(L.lib().dp_attn_bwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(out), L.ptr(g), L.ptr(lse), B, N, ctx.scale,
                                    L.ptr(items), items.shape[0], L.ptr(segs), segs.shape[0], L.ptr(dq), L.ptr(dk),
                                    L.ptr(dv), L.ptr(delta), L.stream()))
        return dq, dk, dv, None, None, None


def attention(q, k, v, hr, wr, ws, scale):
    """q,k,v fp32 [B][N][32] -> fp32 [B][N][32]"""
    items, segs, _ = attention_segments(hr, wr, ws, q.device)
    return _Attention.apply(q.contiguous(), k.contiguous(), v.contiguous(), items, segs, float(scale))


@_internal
def add(a, b, c=None):
    """elementwise a + b (+ c) on dense NHWC bf16 (autograd: plain fan-out)."""
    return _Add.apply(a, b, c)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        a_, b_ = (a if a.is_contiguous() else a.contiguous()), (b if b.is_contiguous() else b.contiguous())
        c_ = None if c is None else (c if c.is_contiguous() else c.contiguous())
        out = torch.empty_like(a_)
        L.check(L.lib().dp_add_bf16(L.ptr(a_), L.ptr(b_), L.ptr(c_), L.ptr(out), out.numel(), L.stream()))
        ctx.has_c = c is not None
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g, (g if ctx.has_c else None)


class _Concat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        B, H, W, Ca = a.shape
        Cb = b.shape[-1]
        out = _nhwc(B, H, W, Ca + Cb, a.device)
        npix = B * H * W
        L.check(L.lib().dp_copy_channels(L.ptr(a), _ld(a), L.ptr(out), Ca + Cb, npix, Ca, L.stream()))
        L.check(L.lib().dp_copy_channels(L.ptr(b), _ld(b), out.data_ptr() + 2 * Ca, Ca + Cb, npix, Cb, L.stream()))
        ctx.split = Ca
        return out

    @staticmethod
    def backward(ctx, g):
        # dense halves (one strided-read copy kernel each): strided views would make every consumer - and autograd's
        # own gradient accumulation - fall back to slow non-vectorised elementwise kernels
        g = g if g.stride(3) == 1 else g.contiguous()
        B, H, W, C = g.shape
        ld = _ld(g)
        npix = B * H * W
        outs = []
        for lo, hi in ((0, ctx.split), (ctx.split, C)):
            t = _nhwc(B, H, W, hi - lo, g.device)
            L.check(L.lib().dp_copy_channels(g.data_ptr() + 2 * lo, ld, L.ptr(t), hi - lo, npix, hi - lo, L.stream()))
            outs.append(t)
        return outs[0], outs[1]


@_internal
def concat_channels(a, b):
    """torch.cat([a, b], dim=1) of the reference, in NHWC."""
    return _Concat.apply(a, b)


class _Relu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xd = x if x.is_contiguous() else x.contiguous()
        out = torch.empty_like(xd)
        L.check(L.lib().dp_relu_bf16(L.ptr(xd), L.ptr(out), out.numel(), L.stream()))
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return _mask_grad(None, g, y)


@_internal
def relu(x):
    return _Relu.apply(x)
