"""Op-level numerics: every hand-written kernel behind ops.py against a plain PyTorch fp32 reference of the same
op, evaluated on the same bf16-rounded operands.  Tolerances are bf16 output rounding (2^-8 relative) plus
accumulation-order slack, stated per test."""
import os
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(BF).float()


def nhwc(x):  # NCHW fp32 -> NHWC bf16 cuda
    return x.permute(0, 2, 3, 1).contiguous().to(BF).cuda()


def nchw(y):  # NHWC bf16 cuda -> NCHW fp32 cpu
    return y.float().permute(0, 3, 1, 2).cpu()


def close(a, b, rel=2 ** -7, abs_frac=4e-3):
    """|a-b| <= rel*|b| + abs_frac*max|b|"""
    tol = rel * b.abs() + abs_frac * float(b.abs().max())
    bad = (a - b).abs() > tol
    return not bool(bad.any()), float((a - b).abs().max()), float(b.abs().max())


def test_layout_roundtrip(pkg):
    from depth_b200 import ops
    x = rnd(2, 136, 7, 9, seed=1).cuda().requires_grad_(True)
    y = ops.to_nhwc(x)
    assert y.shape == (2, 7, 9, 136) and y.dtype == BF
    assert torch.equal(y.float().permute(0, 3, 1, 2), x.detach())
    z = ops.to_nchw(y)
    assert torch.equal(z, x.detach())
    z.backward(torch.ones_like(z) * 0.5)
    assert torch.equal(x.grad, torch.full_like(x, 0.5))


@pytest.mark.parametrize("C", [1, 3, 8])
def test_layout_small_c_into_8_channel_pixels(pkg, C):
    """the stem's input path: C <= 8 planes into 8-channel NHWC pixels, pad channels written as zeros (the destination
    starts as NaN), values rounded to bf16 exactly as the tiled converter does."""
    from depth_b200 import _lib as L
    x = rnd(2, C, 13, 21, seed=11).cuda()
    out = torch.full((2, 13, 21, 8), float("nan"), dtype=BF, device="cuda")
    L.check(L.lib().dp_nchw_f32_to_nhwc_bf16(L.ptr(x), 2, C, 13, 21, L.ptr(out), 8, L.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out[..., :C].float(), x.permute(0, 2, 3, 1).to(BF).float())
    assert bool((out[..., C:] == 0).all())


@pytest.mark.parametrize("B,C,Hi,Wi,Ho,Wo,align", [
    (2, 32, 14, 18, 28, 36, True), (1, 64, 9, 11, 18, 22, False), (2, 16, 16, 20, 224 // 8, 280 // 8, True),
    (1, 8, 28, 35, 56, 72, True), (1, 32, 56, 70, 112, 144, True), (1, 8, 13, 17, 7, 9, True), (1, 8, 6, 6, 6, 6, False),
    (1, 8, 10, 12, 20, 24, False), (1, 256, 14, 18, 28, 36, True), (1, 64, 32, 40, 56, 70, True),
    (2, 128, 24, 30, 48, 60, False), (1, 40, 23, 29, 46, 57, True), (1, 16, 31, 33, 9, 11, False)])
def test_resize_fwd_bwd(pkg, B, C, Hi, Wi, Ho, Wo, align):
    from depth_b200 import ops
    x = rnd(B, C, Hi, Wi, seed=3)
    xr = x.clone().requires_grad_(True)
    ref = F.interpolate(xr, size=(Ho, Wo), mode="bilinear", align_corners=align)
    cot = rnd(B, C, Ho, Wo, seed=4)
    ref.backward(cot)
    xp = nhwc(x).requires_grad_(True)
    out = ops.resize(xp, (Ho, Wo), align)
    if (Ho, Wo) != (Hi, Wi):
        out.backward(nhwc(cot))
        ok, e, s = close(nchw(xp.grad), xr.grad)
        assert ok, ("bwd", e, s)
    ok, e, s = close(nchw(out.detach()), ref.detach())
    assert ok, ("fwd", e, s)


def test_resize_sweep_small_and_strided(pkg):
    """tools/resize_sweep.py: tiny and odd shapes (2x3 .. 32x48), both corner conventions, 8 .. 256 channels, dense and
    channel-slice operands, forward and backward against F.interpolate."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "resize_sweep.py")
    spec = importlib.util.spec_from_file_location("resize_sweep", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.sweep(channels=(8, 32, 256)) == []


def test_resize_planes_f32(pkg):
    from depth_b200 import ops
    x = torch.randn(2, 3, 64, 96).cuda()
    for size, align in [((224, 280), True), ((426, 560), True), ((32, 48), False)]:
        ref = F.interpolate(x, size=size, mode="bilinear", align_corners=align)
        out = ops.resize_planes_f32(x, size, align)
        assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)


CONV_CASES = [
    # kind, Cin, Cout, k, stride, pad, B, H, W
    ("conv", 32, 32, 3, 2, 1, 2, 32, 48), ("conv", 32, 32, 3, 2, 1, 1, 17, 23), ("conv", 512, 512, 3, 2, 1, 1, 16, 20),
    ("convT", 32, 32, 4, 2, 1, 2, 8, 12), ("convT", 128, 128, 4, 4, 0, 1, 4, 5), ("convT", 256, 256, 2, 2, 0, 1, 4, 5),
    ("conv", 32, 32, 3, 2, 1, 2, 224, 288), ("convT", 32, 32, 4, 2, 1, 2, 112, 144), ("conv", 64, 48, 3, 2, 1, 1, 18, 30),
    ("convT", 64, 32, 4, 2, 1, 1, 9, 15), ("conv", 32, 32, 3, 2, 1, 1, 8, 8), ("convT", 16, 16, 4, 2, 1, 1, 7, 9),
]


@pytest.mark.parametrize("kind,Cin,Cout,k,stride,pad,B,H,W", CONV_CASES)
def test_strided_and_transposed_conv(pkg, kind, Cin, Cout, k, stride, pad, B, H, W):
    from depth_b200 import ops
    x = rnd(B, Cin, H, W, seed=5)
    if kind == "conv":
        m = nn.Conv2d(Cin, Cout, k, stride, pad)
    else:
        m = nn.ConvTranspose2d(Cin, Cout, k, stride, pad)
    with torch.no_grad():
        m.weight.copy_(rnd(*m.weight.shape, seed=6, scale=(1.0 / (Cin * k * k / stride ** 2)) ** 0.5))
        m.bias.copy_(rnd(Cout, seed=7, scale=0.1))
    xr = x.clone().requires_grad_(True)
    ref = m(xr)
    cot = rnd(*ref.shape, seed=8)
    ref.backward(cot)
    mc = type(m)(Cin, Cout, k, stride, pad).cuda()
    mc.load_state_dict(m.state_dict())
    xp = nhwc(x).requires_grad_(True)
    fn = ops.conv_strided if kind == "conv" else ops.conv_transposed
    out = fn(xp, mc.weight, mc.bias, stride, pad)
    assert tuple(out.shape) == (B, ref.shape[2], ref.shape[3], Cout)
    out.backward(nhwc(cot))
    for nm, a, b in [("fwd", nchw(out.detach()), ref.detach()), ("dx", nchw(xp.grad), xr.grad),
                     ("dw", mc.weight.grad.cpu(), m.weight.grad), ("db", mc.bias.grad.cpu(), m.bias.grad)]:
        ok, e, s = close(a, b)
        assert ok, (nm, e, s)


@pytest.mark.parametrize("H,W", [(20, 28), (11, 37), (3, 5)])       # ragged strips / single-row images for the 16 -> 1 fast path
@pytest.mark.parametrize("C,KS,relu", [(16, 3, True), (16, 3, False), (32, 1, True), (32, 1, False), (64, 3, True)])
def test_head_conv(pkg, C, KS, relu, H, W):
    from depth_b200 import ops
    B = 3
    x = rnd(B, C, H, W, seed=9)
    m = nn.Conv2d(C, 1, KS, 1, KS // 2)
    with torch.no_grad():
        m.bias.fill_(0.1)
    xr = x.clone().requires_grad_(True)
    ref = m(xr)
    if relu:
        ref = F.relu(ref)
    cot = torch.rand(B, 1, H, W) + 0.5
    ref.backward(cot)
    mc = nn.Conv2d(C, 1, KS, 1, KS // 2).cuda()
    mc.load_state_dict(m.state_dict())
    xp = nhwc(x).requires_grad_(True)
    out = ops.head_conv(xp, mc.weight, mc.bias, relu)
    assert out.shape == (B, H, W) and out.dtype == torch.float32
    out.backward(cot[:, 0].cuda())
    assert torch.allclose(out.detach().cpu(), ref.detach()[:, 0], rtol=1e-4, atol=1e-5)
    ok, e, s = close(nchw(xp.grad), xr.grad)
    assert ok, ("dx", e, s)
    assert torch.allclose(mc.weight.grad.cpu(), m.weight.grad, rtol=1e-3, atol=1e-3 * float(m.weight.grad.abs().max()))
    assert torch.allclose(mc.bias.grad.cpu(), m.bias.grad, rtol=1e-4)


@pytest.mark.parametrize("C,mode", [(64, "plain"), (32, "res"), (16, "two"), (32, "eval"), (32, "norelu")])
def test_batchnorm_act(pkg, C, mode):
    """train / eval BatchNorm (+residual | + second BN branch) + ReLU, forward, backward and running statistics.
    The backward is compared at the scale of the incoming gradient (its own cancellation is the op's nature)."""
    from depth_b200 import ops
    B, H, W = 2, 24, 40
    x = rnd(B, C, H, W, seed=11) * 1.5 + 0.3
    x = x.to(BF).float()
    x2 = rnd(B, C, H, W, seed=12).to(BF).float()
    bn, bn2 = nn.BatchNorm2d(C), nn.BatchNorm2d(C)
    with torch.no_grad():
        for b_ in (bn, bn2):
            b_.weight.copy_(1 + 0.2 * torch.randn(C)); b_.bias.copy_(0.1 * torch.randn(C))
            b_.running_mean.copy_(0.1 * torch.randn(C)); b_.running_var.copy_(1 + 0.1 * torch.rand(C))
    bnc, bn2c = nn.BatchNorm2d(C).cuda(), nn.BatchNorm2d(C).cuda()
    bnc.load_state_dict(bn.state_dict()); bn2c.load_state_dict(bn2.state_dict())
    if mode == "eval":
        bn.eval(); bnc.eval()
    xr, x2r = x.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    if mode == "two":
        ref = F.relu(bn(xr) + bn2(x2r))
    elif mode == "res":
        ref = F.relu(bn(xr) + x2r)
    elif mode == "norelu":
        ref = bn(xr)
    else:
        ref = F.relu(bn(xr))
    cot = (torch.rand(B, C, H, W) + 0.5).to(BF).float()
    ref.backward(cot)
    xp, x2p = nhwc(x).requires_grad_(True), nhwc(x2).requires_grad_(True)
    if mode == "two":
        out = ops.bn_act(bnc, xp, relu=True, bn2=bn2c, c2=x2p)
    elif mode == "res":
        out = ops.bn_act(bnc, xp, relu=True, res=x2p)
    else:
        out = ops.bn_act(bnc, xp, relu=(mode != "norelu"))
    out.backward(nhwc(cot))
    ok, e, s = close(nchw(out.detach()), ref.detach())
    assert ok, ("fwd", e, s)
    gscale = float(cot.abs().max()) * float(bn.weight.detach().abs().max()) * 2
    assert float((nchw(xp.grad) - xr.grad).abs().max()) < 2e-2 * gscale
    if mode in ("two", "res"):
        assert float((nchw(x2p.grad) - x2r.grad).abs().max()) < 2e-2 * gscale
    n = B * H * W
    assert torch.allclose(bnc.weight.grad.cpu(), bn.weight.grad, rtol=2e-2, atol=2e-3 * n)
    assert torch.allclose(bnc.bias.grad.cpu(), bn.bias.grad, rtol=2e-2, atol=2e-3 * n)
    if mode == "two":
        assert torch.allclose(bn2c.weight.grad.cpu(), bn2.weight.grad, rtol=2e-2, atol=2e-3 * n)
    assert torch.allclose(bnc.running_mean.cpu(), bn.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bnc.running_var.cpu(), bn.running_var, rtol=1e-4, atol=1e-5)
    assert int(bnc.num_batches_tracked) == int(bn.num_batches_tracked)


def test_ln_linear(pkg):
    from depth_b200 import ops
    B, N, D = 2, 300, 32
    ln, lin = nn.LayerNorm(D), nn.Linear(D, D, bias=True)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.2 * torch.randn(D)); ln.bias.copy_(0.1 * torch.randn(D))
    lnc, linc = nn.LayerNorm(D).cuda(), nn.Linear(D, D, bias=True).cuda()
    lnc.load_state_dict(ln.state_dict()); linc.load_state_dict(lin.state_dict())
    for in_bf16, out_bf16 in [(True, False), (False, True), (False, False)]:
        for m in (ln, lin, lnc, linc):
            m.zero_grad()
        x = rnd(B, N, D, seed=13) if in_bf16 else torch.randn(B, N, D)
        xr = x.clone().requires_grad_(True)
        ref = lin(ln(xr))
        cot = torch.randn(B, N, D).to(BF).float() if out_bf16 else torch.randn(B, N, D)
        ref.backward(cot)
        xp = (x.to(BF) if in_bf16 else x).cuda().requires_grad_(True)
        out = ops.ln_linear(xp, lnc, linc, out_bf16=out_bf16)
        out.backward(cot.to(out.dtype).cuda())
        tol = 1e-2 if (in_bf16 or out_bf16) else 2e-4
        assert torch.allclose(out.detach().float().cpu(), ref.detach(), rtol=tol, atol=tol)
        assert torch.allclose(xp.grad.float().cpu(), xr.grad, rtol=tol, atol=tol)
        for a, b in [(linc.weight.grad, lin.weight.grad), (linc.bias.grad, lin.bias.grad),
                     (lnc.weight.grad, ln.weight.grad), (lnc.bias.grad, ln.bias.grad)]:
            assert torch.allclose(a.cpu(), b, rtol=1e-3, atol=1e-3 * float(b.abs().max()) + 1e-4)


@pytest.mark.parametrize("hr,wr", [(8, 12), (20, 24), (56, 72), (17, 40)])
def test_segmented_attention_vs_window_loop(pkg, hr, wr):
    """the last-writer kernel against the reference's own overwrite loop (midas_semantics.py:93-112), fwd + bwd"""
    from depth_b200 import ops
    from oracle.model import CrossAttention
    B, N, D, nh, hd, ws = 2, hr * wr, 32, 8, 4, 16
    g = torch.Generator().manual_seed(21)
    q0, k0, v0 = (torch.randn(B, N, D, generator=g) for _ in range(3))
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q0, k0, v0))
    q, k, v = (t.reshape(B, N, nh, hd).permute(0, 2, 1, 3) for t in (qr, kr, vr))
    out = torch.zeros(B, N, D)
    for lo, hi in CrossAttention.window_ranges(hr, wr, ws):
        a = ((q[:, :, lo:hi] @ k[:, :, lo:hi].transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
        out[:, lo:hi, :] = (a @ v[:, :, lo:hi]).transpose(1, 2).reshape(B, -1, D)
    cot = torch.randn(B, N, D, generator=g)
    out.backward(cot)
    qc, kc, vc = (t.clone().cuda().requires_grad_(True) for t in (q0, k0, v0))
    oc = ops.attention(qc, kc, vc, hr, wr, ws, hd ** -0.5)
    oc.backward(cot.cuda())
    assert torch.allclose(oc.detach().cpu(), out.detach(), rtol=1e-4, atol=1e-5)
    for a, b in [(qc.grad, qr.grad), (kc.grad, kr.grad), (vc.grad, vr.grad)]:
        assert torch.allclose(a.cpu(), b, rtol=1e-3, atol=1e-4 * float(b.abs().max()) + 1e-6)


def test_concat_add_relu(pkg):
    from depth_b200 import ops
    a, b = rnd(2, 32, 6, 10, seed=30), rnd(2, 32, 6, 10, seed=31)
    ap, bp = nhwc(a).requires_grad_(True), nhwc(b).requires_grad_(True)
    cat = ops.concat_channels(ap, bp)
    assert torch.equal(nchw(cat.detach()), torch.cat([a, b], 1))
    s = ops.add(ops.relu(cat[..., :32].contiguous()), bp)
    cot = nhwc(rnd(2, 32, 6, 10, seed=32))
    s.backward(cot)
    assert torch.allclose(nchw(s.detach()), (F.relu(a) + b).to(BF).float(), atol=1e-6)
    assert torch.equal(nchw(ap.grad), (nchw(cot) * (a > 0)).to(BF).float())
    assert torch.equal(nchw(bp.grad), nchw(cot))
