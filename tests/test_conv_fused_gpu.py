"""BatchNorm fused around the tcgen05 convolutions (dp_conv2d_tc_fused / dp_conv2d_wgrad_tc_fused), through the C ABI.

The fused launches must reproduce the unfused pipeline (dp_bn_apply -> dp_conv2d_tc, dp_conv2d_tc -> dp_chan_reduce) BIT FOR
BIT: the prologue computes act(fma(c, scale, shift)) with the same fp32 arithmetic and the same single bf16 rounding as
dp_bn_apply, and keeps the zero padding of the ACTIVATED tensor (reference: nn.Conv2d(padding=1) after nn.ReLU,
midas_semantics.py:145-147).  They are also checked against a plain PyTorch fp32 reference of the same op."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from tests.test_conv_tc_gpu import pack_w

pytestmark = pytest.mark.gpu


def _ss(C, seed, positive_bias=0.0):
    g = torch.Generator().manual_seed(seed)
    scale = (torch.rand(C, generator=g) * 1.5 + 0.25) * torch.where(torch.rand(C, generator=g) < 0.2, -1.0, 1.0)
    shift = torch.randn(C, generator=g) * 0.5 + positive_bias
    return torch.stack([scale, shift]).contiguous().cuda()


def _bn_apply(L, c, ss, act):
    B, H, W, C = c.shape
    y = torch.empty_like(c)
    L.check(L.lib().dp_bn_apply(L.ptr(c), C, L.ptr(ss), None, 0, None, None, 0, B * H * W, C, act, L.ptr(y), C, L.stream()))
    return y


def _conv(L, x, wp, Cout, KS, fuse=None, res=None, stats=None):
    B, H, W, Cin = x.shape
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib().dp_conv2d_tc_fused(L.ptr(x), Cin, B, H, W, Cin, L.ptr(wp), wp.shape[2], Cout, KS, None,
                                       L.ptr(res), Cout, None, 0, 0, L.ptr(out), Cout, None, 0, 0, L.ptr(stats),
                                       ctypes.byref(fuse) if fuse is not None else None, L.stream()))
    return out


PRE_CASES = [
    # B, H, W, Cin, Cout, KS, act
    (2, 45, 52, 64, 64, 3, 1),      # halo mode, 128-byte rows, ragged borders
    (2, 37, 43, 32, 32, 3, 1),      # halo, 64-byte rows, per-warp epilogue
    (1, 50, 70, 16, 16, 3, 1),      # halo, 32-byte rows
    (2, 24, 40, 32, 16, 3, 1),
    (1, 33, 47, 64, 32, 1, 1),      # 1x1 shortcut
    (1, 28, 36, 128, 128, 3, 1),    # three column loads, streamed weights, two k-chunks
    (2, 56, 72, 144, 32, 1, 2),     # trunk projection: ReLU6, Cin not a multiple of the 64-channel chunk
    (2, 28, 36, 576, 96, 1, 2),
    (1, 14, 18, 1392, 232, 1, 2),
    (1, 448, 576, 64, 64, 3, 1),    # the benched shape: many tiles per persistent CTA
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,KS,act", PRE_CASES)
def test_prologue_matches_unfused_bitwise(pkg, B, H, W, Cin, Cout, KS, act):
    L = pkg._lib
    g = torch.Generator().manual_seed(B * 31 + H + Cin + Cout)
    c = (torch.randn(B, H, W, Cin, generator=g) * 2).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, KS, KS, generator=g) * (2.0 / (Cin * KS * KS)) ** 0.5).cuda()
    wp = pack_w(w)
    ss = _ss(Cin, Cin + KS, positive_bias=2.0 if act == 2 else 0.0)
    assert L.lib().dp_conv2d_tc_caps(B, H, W, Cin, Cout, KS) & L.CAP_PROLOGUE
    a = _bn_apply(L, c, ss, act)
    ref_out = _conv(L, a, wp, Cout, KS)
    fuse = L.ConvFuse(L.ptr(ss), act, None, 0, None, 0)
    gsz = L.lib().dp_conv2d_tc_grid(B, H, W, Cin, Cout, KS)
    st = torch.empty(gsz, 2, Cout, device="cuda")
    st_ref = torch.empty_like(st)
    out = _conv(L, c, wp, Cout, KS, fuse=fuse, stats=st)
    _conv(L, a, wp, Cout, KS, stats=st_ref)
    torch.cuda.synchronize()
    assert torch.equal(out, ref_out), float((out.float() - ref_out.float()).abs().max())
    assert torch.allclose(st.sum(0), st_ref.sum(0), rtol=1e-5, atol=1e-3)
    # and against PyTorch fp32 on the same bf16-rounded activation
    y = F.conv2d(a.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), None, padding=KS // 2).permute(0, 2, 3, 1)
    err = (out.float() - y).abs()
    assert bool((err <= 2 ** -7 * y.abs() + 4e-3).all()), float(err.max())


MASK_CASES = [(2, 45, 52, 64, 64, 3, 1), (2, 37, 43, 32, 32, 3, 1), (1, 50, 70, 16, 16, 3, 1), (2, 24, 40, 16, 32, 3, 1),
              (2, 33, 47, 16, 32, 1, 1), (1, 30, 44, 32, 64, 3, 2)]


@pytest.mark.parametrize("B,H,W,Cin,Cout,KS,act", MASK_CASES)
def test_bn_backward_epilogue_matches_reduce_pass(pkg, B, H, W, Cin, Cout, KS, act):
    """a data-gradient launch (Cin = channels of dY, Cout = channels of the gradient it produces) with the mask epilogue
    vs the same launch followed by dp_chan_reduce(mode 2)."""
    L = pkg._lib
    lib = L.lib()
    g = torch.Generator().manual_seed(H * 3 + Cin + Cout)
    dy = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, KS, KS, generator=g) * (2.0 / (Cin * KS * KS)) ** 0.5).cuda()
    wp = pack_w(w)
    c = (torch.randn(B, H, W, Cout, generator=g) * 2).to(torch.bfloat16).cuda()
    ss = _ss(Cout, Cout, positive_bias=2.0 if act == 2 else 0.0)
    assert lib.dp_conv2d_tc_caps(B, H, W, Cin, Cout, KS) & L.CAP_BN_BACKWARD
    npix = B * H * W
    # unfused: raw gradient, then the masked reduction pass
    graw = _conv(L, dy, wp, Cout, KS)
    nb = lib.dp_chan_reduce_blocks()
    part = torch.empty(nb, 2, Cout, device="cuda")
    L.check(lib.dp_chan_reduce(3 if act == 2 else 2, L.ptr(c), Cout, L.ptr(graw), Cout, None, 0, L.ptr(ss), npix, Cout,
                               L.ptr(part), L.stream()))
    m = c.float() * ss[0] + ss[1]
    keep = (m > 0) & ((m < 6) if act == 2 else torch.ones_like(m, dtype=torch.bool))
    gref = torch.where(keep, graw.float(), torch.zeros_like(m)).to(torch.bfloat16)
    # fused
    gsz = lib.dp_conv2d_tc_grid(B, H, W, Cin, Cout, KS)
    st = torch.empty(gsz, 2, Cout, device="cuda")
    fuse = L.ConvFuse(None, 0, L.ptr(c), Cout, L.ptr(ss), act)
    gout = _conv(L, dy, wp, Cout, KS, fuse=fuse, stats=st)
    torch.cuda.synchronize()
    assert torch.equal(gout, gref), float((gout.float() - gref.float()).abs().max())
    got, want = st.double().sum(0), part.double().sum(0)
    scale = want.abs().max(dim=1, keepdim=True).values + 1e-6
    assert bool(((got - want).abs() <= 1e-4 * scale + 1e-3).all()), (got - want).abs().max()


WG_PRE = [(2, 45, 52, 64, 64, 3, 1), (2, 37, 43, 32, 32, 3, 1), (1, 50, 70, 16, 16, 3, 1), (2, 24, 40, 32, 16, 3, 1),
          (1, 33, 47, 64, 32, 1, 1), (1, 28, 36, 128, 128, 3, 1), (2, 56, 72, 144, 32, 1, 2), (1, 14, 18, 1392, 232, 1, 2),
          (1, 224, 288, 64, 64, 3, 1)]


@pytest.mark.parametrize("B,H,W,Cin,Cout,KS,act", WG_PRE)
def test_wgrad_prologue_matches_unfused_bitwise(pkg, B, H, W, Cin, Cout, KS, act):
    L = pkg._lib
    lib = L.lib()
    g = torch.Generator().manual_seed(H + 5 * Cin + Cout)
    c = (torch.randn(B, H, W, Cin, generator=g) * 2).to(torch.bfloat16).cuda()
    dy = torch.randn(B, H, W, Cout, generator=g).to(torch.bfloat16).cuda()
    ss = _ss(Cin, Cin * 3 + 1, positive_bias=2.0 if act == 2 else 0.0)
    a = _bn_apply(L, c, ss, act)
    nb = lib.dp_conv2d_wgrad_tc_workspace(B, H, W, Cin, Cout, KS)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    ref = torch.full((Cout, Cin, KS, KS), float("nan"), device="cuda")
    got = torch.full_like(ref, float("nan"))
    L.check(lib.dp_conv2d_wgrad_tc(L.ptr(a), Cin, L.ptr(dy), Cout, B, H, W, Cin, Cout, KS, L.ptr(ref), 0, L.ptr(ws), nb,
                                   L.stream()))
    L.check(lib.dp_conv2d_wgrad_tc_fused(L.ptr(c), Cin, L.ptr(dy), Cout, B, H, W, Cin, Cout, KS, L.ptr(got), 0, L.ptr(ws),
                                         nb, L.ptr(ss), act, L.stream()))
    torch.cuda.synchronize()
    assert torch.equal(got, ref), float((got - ref).abs().max())
