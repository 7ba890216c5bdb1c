"""Stress test of the two-issuer tcgen05 pipeline (csrc/conv_tc.cu, make_plan: ring depth = multiple of 2 * k-steps).

Round 1 parked an intermittent fault of the two-issuer mode with several k-steps per tile (EfficientNet-Lite3 trunk 1x1
layers, Cin 144 / 192).  Root cause: mbarrier parity aliasing - an issuer that skips the other issuer's laps of a shared
full[] barrier can be answered "complete" by the lap two phases back.  The ring is now cut so that every barrier has a
single producer and a single consumer; this test hammers exactly those shapes (and the benched full-resolution ones)
with a second stream generating memory traffic, and demands bit-identical results on every launch plus agreement with
the fp32 reference.  (compute-sanitizer is closed on this GPU pool, so the evidence is this test.)"""
import pytest
import torch

from tests.test_conv_tc_gpu import pack_w, ref_conv

pytestmark = pytest.mark.gpu

# B, H, W, Cin, Cout, KS, launches
SHAPES = [
    (4, 112, 144, 144, 32, 1, 150),    # three k-steps, BN = 32: two issuers, ring of 6
    (4, 112, 144, 192, 32, 1, 150),
    (4, 56, 72, 192, 48, 1, 150),
    (4, 56, 72, 288, 48, 1, 100),      # five k-steps: falls back to one issuer
    (2, 112, 144, 128, 64, 1, 100),    # two k-steps: ring of 8
    (2, 224, 288, 32, 32, 3, 100),     # halo mode, single k-step, per-warp epilogue
    (1, 448, 576, 64, 64, 3, 60),      # the benched dominant shape
    (1, 448, 576, 16, 16, 3, 60),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,KS,n", SHAPES)
def test_repeated_launches_under_load_are_bit_identical(pkg, B, H, W, Cin, Cout, KS, n):
    L = pkg._lib
    g = torch.Generator().manual_seed(Cin * 7 + Cout + KS)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, KS, KS, generator=g) * (2.0 / (Cin * KS * KS)) ** 0.5).cuda()
    wp = pack_w(w)
    ref = ref_conv(x, w, None, None, KS)
    outs = [torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
    gsz = L.lib().dp_conv2d_tc_grid(B, H, W, Cin, Cout, KS)
    stats = [torch.empty(gsz, 2, Cout, device="cuda") for _ in range(2)]
    noise_a = torch.empty(64 << 20, device="cuda", dtype=torch.uint8)
    noise_b = torch.empty_like(noise_a)
    side = torch.cuda.Stream()
    first = None
    for i in range(n):
        if i % 3 == 0:
            with torch.cuda.stream(side):      # HBM / L2 pressure from a second stream while the conv runs
                noise_b.copy_(noise_a)
        o, st = outs[i & 1], stats[i & 1]
        o.fill_(float("nan"))
        L.check(L.lib().dp_conv2d_tc(L.ptr(x), Cin, B, H, W, Cin, L.ptr(wp), wp.shape[2], Cout, KS, None, None, 0, None,
                                     0, 0, L.ptr(o), Cout, None, 0, 0, L.ptr(st), L.stream()))
        if first is None:
            torch.cuda.synchronize()
            first = (o.clone(), st.clone())
            err = (o.float() - ref).abs()
            assert bool((err <= 2 ** -7 * ref.abs() + 2e-3).all()), float(err.max())
        elif i % 10 == 9 or i == n - 1:
            torch.cuda.synchronize()
            assert torch.equal(o, first[0]), f"launch {i} differs from launch 0"
            assert torch.equal(st, first[1]), f"BN partials of launch {i} differ"
    torch.cuda.synchronize()
