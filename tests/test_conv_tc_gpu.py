"""tcgen05 implicit-GEMM convolution vs a plain PyTorch fp32 reference on bf16-rounded operands.
Tolerance: fp32 accumulation of exactly representable products -> only the final bf16 rounding of the
output differs: |err| <= 2^-8 * |ref| + small absolute slack for accumulation order."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def pack_w(w):
    """OIHW fp32 -> [KS*KS][Cout][Cin] bf16"""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous().to(torch.bfloat16)


def run_conv(pkg, x_nhwc, wp, Cout, KS, bias=None, res=None, relu=False, out2=False, relu2=False, stats=False):
    L = pkg._lib
    B, H, W, Cin = x_nhwc.shape
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty_like(out) if out2 else None
    st = None
    if stats:
        g = L.lib().dp_conv2d_tc_grid(B, H, W, Cin, Cout, KS)
        st = torch.full((g, 2, Cout), float("nan"), device="cuda", dtype=torch.float32)
    L.check(L.lib().dp_conv2d_tc(L.ptr(x_nhwc), Cin, B, H, W, Cin, L.ptr(wp), wp.shape[2], Cout, KS, L.ptr(bias),
                                 L.ptr(res), Cout, None, 0, int(relu), L.ptr(out), Cout, L.ptr(o2), Cout, int(relu2),
                                 L.ptr(st), L.stream()))
    torch.cuda.synchronize()
    return out, o2, st


def ref_conv(x_nhwc, w, bias, res, KS):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w.to(torch.bfloat16).float(), bias, padding=KS // 2)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + res.float()
    return y


CASES = [
    # B, H, W, Cin, Cout, KS
    (2, 16, 32, 64, 64, 3),
    (1, 14, 18, 512, 512, 3),
    (2, 28, 36, 256, 256, 3),
    (1, 56, 72, 128, 128, 3),
    (2, 31, 45, 64, 32, 3),
    (2, 24, 40, 32, 32, 3),
    (1, 24, 40, 32, 16, 3),
    (1, 16, 24, 16, 32, 3),
    (1, 28, 36, 136, 256, 3),
    (1, 56, 72, 48, 128, 3),
    (1, 28, 36, 256, 136, 3),
    (1, 14, 18, 512, 384, 3),
    (2, 20, 28, 512, 256, 1),
    (1, 33, 47, 64, 64, 1),
    (1, 24, 40, 64, 32, 1),
    (1, 16, 20, 384, 128, 1),
    # EfficientNet-Lite3 trunk 1x1 shapes: many N blocks, Cout / Cin that are not multiples of 16 / 64
    (1, 14, 18, 384, 1392, 1), (1, 14, 18, 1392, 232, 1), (2, 14, 18, 232, 1392, 1), (1, 28, 36, 816, 136, 1),
    (1, 56, 72, 24, 144, 1), (1, 56, 72, 144, 24, 1), (1, 28, 36, 96, 576, 1), (1, 14, 18, 1392, 384, 1),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,KS", CASES)
def test_conv_plain(pkg, B, H, W, Cin, Cout, KS):
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, KS, KS, generator=g) * (2.0 / (Cin * KS * KS)) ** 0.5).cuda()
    out, _, _ = run_conv(pkg, x, pack_w(w), Cout, KS)
    ref = ref_conv(x, w, None, None, KS)
    err = (out.float() - ref).abs()
    tol = 2 ** -7 * ref.abs() + 2e-3
    assert bool((err <= tol).all()), f"max err {float(err.max())} at ref scale {float(ref.abs().max())}"


def test_conv_epilogue_bias_residual_relu_dual(pkg):
    B, H, W, Cin, Cout = 2, 20, 24, 64, 64
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.06).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(B, H, W, Cout, generator=g).to(torch.bfloat16).cuda()
    out, o2, _ = run_conv(pkg, x, pack_w(w), Cout, 3, bias=bias, res=res, relu=False, out2=True, relu2=True)
    ref = ref_conv(x, w, bias, res, 3)
    tol = 2 ** -7 * ref.abs() + 2e-3
    assert bool(((out.float() - ref).abs() <= tol).all())
    assert bool(((o2.float() - ref.clamp_min(0)).abs() <= tol).all())
    out, _, _ = run_conv(pkg, x, pack_w(w), Cout, 3, bias=bias, relu=True)
    ref = ref_conv(x, w, bias, None, 3).clamp_min(0)
    assert bool(((out.float() - ref).abs() <= 2 ** -7 * ref.abs() + 2e-3).all())


@pytest.mark.parametrize("Cin,Cout,KS", [(64, 64, 3), (32, 32, 3), (64, 32, 1), (32, 16, 3), (16, 16, 3), (16, 8, 1)])
def test_conv_tma_store_epilogue_all_operands(pkg, Cin, Cout, KS):
    """BN in {16,32,64}: registers -> swizzled smem tile -> TMA store.  Bias + two residuals + dual (raw / ReLU) outputs,
    tile overhang in both directions, outputs and residuals living in channel slices of wider buffers."""
    L = pkg._lib
    B, H, W = 2, 37, 43
    g = torch.Generator().manual_seed(Cin * 100 + Cout + KS)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, KS, KS, generator=g) * (2.0 / (Cin * KS * KS)) ** 0.5).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    wide = Cout + 24
    r1 = torch.randn(B, H, W, wide, generator=g).to(torch.bfloat16).cuda()
    r2 = torch.randn(B, H, W, Cout, generator=g).to(torch.bfloat16).cuda()
    ob = torch.full((B, H, W, wide), 7.0, device="cuda", dtype=torch.bfloat16)
    o2 = torch.full((B, H, W, Cout), 7.0, device="cuda", dtype=torch.bfloat16)
    res1 = r1[..., 8:8 + Cout]
    out = ob[..., 16:16 + Cout]
    L.check(L.lib().dp_conv2d_tc(L.ptr(x), Cin, B, H, W, Cin, L.ptr(pack_w(w)), Cin, Cout, KS, L.ptr(bias),
                                 res1.data_ptr(), wide, L.ptr(r2), Cout, 0, out.data_ptr(), wide, L.ptr(o2), Cout, 1,
                                 None, L.stream()))
    torch.cuda.synchronize()
    ref = ref_conv(x, w, bias, res1, KS) + r2.float()
    tol = 2 ** -7 * ref.abs() + 4e-3
    assert bool(((out.float() - ref).abs() <= tol).all()), float((out.float() - ref).abs().max())
    assert bool(((o2.float() - ref.clamp_min(0)).abs() <= tol).all())
    # the store must not touch the neighbouring channels of the wide buffer
    assert bool((ob[..., :16] == 7.0).all()) and bool((ob[..., 16 + Cout:] == 7.0).all())


@pytest.mark.parametrize("Cin,Cout,KS", [(32, 32, 3), (16, 16, 3), (64, 32, 1), (32, 16, 1), (32, 192, 1), (96, 576, 1),
                                         (48, 288, 1), (136, 816, 1), (24, 144, 1), (128, 128, 3), (64, 128, 3)])
def test_conv_bn_statistics_with_bias_small_n(pkg, Cin, Cout, KS):
    B, H, W = 3, 45, 52
    g = torch.Generator().manual_seed(Cin + 3 * Cout + KS)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, KS, KS, generator=g) * 0.1).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    out, _, st = run_conv(pkg, x, pack_w(w), Cout, KS, bias=bias, stats=True)
    ref = ref_conv(x, w, bias, None, KS)
    assert bool(((out.float() - ref).abs() <= 2 ** -7 * ref.abs() + 2e-3).all())
    s = st.double().sum(dim=0)
    y = out.double().reshape(-1, Cout)
    assert torch.allclose(s[0], y.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[1], (y * y).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("Cin,Cout", [(96, 576), (136, 816), (232, 1392), (32, 192)])
def test_conv_bn_statistics_wide_many_tiles(pkg, Cin, Cout):
    """wide (64-column sub-block) epilogue with several tiles per persistent CTA: the per-N-block sums are folded into
    the CTA's row of the partials whenever the N block changes (3 and 6 N blocks: every tile; 1 and 4: once), and the
    last N block of 816 / 1392 holds sub-blocks without real channels."""
    B, H, W = 8, 45, 52
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, 1, 1, generator=g) * 0.1).cuda()
    out, _, st = run_conv(pkg, x, pack_w(w), Cout, 1, stats=True)
    ref = ref_conv(x, w, None, None, 1)
    assert bool(((out.float() - ref).abs() <= 2 ** -7 * ref.abs() + 2e-3).all())
    s = st.double().sum(dim=0)
    y = out.double().reshape(-1, Cout)
    assert torch.allclose(s[0], y.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[1], (y * y).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 40, 48, 64, 64), (3, 24, 40, 64, 32), (1, 50, 70, 32, 16)])
def test_conv_bn_statistics(pkg, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.06).cuda()
    out, _, st = run_conv(pkg, x, pack_w(w), Cout, 3, stats=True)
    s = st.double().sum(dim=0)
    y = out.double().reshape(-1, Cout)
    assert torch.allclose(s[0], y.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[1], (y * y).sum(0), rtol=1e-4, atol=1e-2)


def test_conv_full_resolution_persistent(pkg):
    """bench-size spatial extent: many tiles per persistent CTA (pipeline phase wrap-around, TMEM double buffering)."""
    B, H, W, Cin, Cout = 2, 448, 576, 64, 64
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.06).cuda()
    out, _, _ = run_conv(pkg, x, pack_w(w), Cout, 3)
    ref = ref_conv(x, w, None, None, 3)
    err = (out.float() - ref).abs()
    assert bool((err <= 2 ** -7 * ref.abs() + 2e-3).all()), float(err.max())
    # linearity (size-independent property): conv(2x) == 2 conv(x) exactly in bf16
    out2, _, _ = run_conv(pkg, (x.float() * 2).to(torch.bfloat16), pack_w(w), Cout, 3)
    assert torch.equal(out2.float(), out.float() * 2)


WG_CASES = [
    (2, 16, 32, 64, 64, 3), (1, 14, 18, 512, 512, 3), (2, 28, 36, 256, 256, 3), (2, 31, 45, 64, 32, 3),
    (2, 24, 40, 32, 32, 3), (1, 24, 40, 32, 16, 3), (1, 24, 40, 16, 16, 3), (1, 28, 36, 136, 256, 3),
    (1, 56, 72, 48, 128, 3), (1, 14, 18, 384, 512, 3), (2, 20, 28, 512, 256, 1), (1, 24, 40, 64, 32, 1),
    (1, 24, 40, 32, 16, 1), (1, 16, 20, 384, 128, 1), (2, 112, 144, 64, 64, 3), (1, 20, 28, 64, 48, 3),
    (1, 20, 28, 32, 64, 3), (1, 20, 28, 128, 136, 3), (1, 20, 28, 48, 48, 1),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,KS", WG_CASES)
def test_wgrad(pkg, B, H, W, Cin, Cout, KS):
    L = pkg._lib
    g = torch.Generator().manual_seed(H * 7 + Cin + Cout)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    dy = torch.randn(B, H, W, Cout, generator=g).to(torch.bfloat16).cuda()
    grad = torch.full((Cout, Cin, KS, KS), float("nan"), device="cuda")
    nb = L.lib().dp_conv2d_wgrad_tc_workspace(B, H, W, Cin, Cout, KS)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    L.check(L.lib().dp_conv2d_wgrad_tc(L.ptr(x), Cin, L.ptr(dy), Cout, B, H, W, Cin, Cout, KS, L.ptr(grad), 0,
                                       L.ptr(ws), nb, L.stream()))
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    w0 = torch.zeros(Cout, Cin, KS, KS, device="cuda", requires_grad=True)
    y = F.conv2d(xr, w0, padding=KS // 2)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = w0.grad
    err = (grad - ref).abs()
    tol = 1e-3 * ref.abs() + 1e-3 * float(ref.abs().max())
    assert bool((err <= tol).all()), f"max err {float(err.max())} scale {float(ref.abs().max())}"
