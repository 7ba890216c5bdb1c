"""EfficientNet-Lite3 trunk kernels: depthwise conv forward / data gradient / weight gradient with
fused BatchNorm statistics, BatchNorm + ReLU6, the 3-channel stem, and the whole trunk against the same hub-shaped
module run by PyTorch in fp32.  Tolerances: bf16 storage of every activation (2^-8 relative per tensor) - stated per test."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(BF).float()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(BF).cuda()


def nchw(y):
    return y.float().permute(0, 3, 1, 2).cpu()


def close(a, b, rel=2 ** -7, abs_frac=4e-3):
    tol = rel * b.abs() + abs_frac * float(b.abs().max())
    return not bool(((a - b).abs() > tol).any()), float((a - b).abs().max()), float(b.abs().max())


@pytest.mark.parametrize("B,C,H,W,K,S,pad", [
    (2, 32, 20, 28, 3, 1, 1), (2, 144, 31, 45, 3, 2, 1), (1, 192, 28, 37, 5, 2, 2), (1, 816, 14, 18, 5, 1, 2),
    (1, 1392, 7, 9, 3, 1, 1), (2, 48, 16, 19, 5, 1, 2), (1, 24, 33, 18, 3, 2, 0),
    # several tiles per persistent block (both tile buffers, both barrier phases), ragged tile borders
    (4, 288, 56, 72, 5, 1, 2), (4, 192, 61, 75, 3, 1, 1), (8, 40, 45, 37, 5, 1, 2)])
def test_depthwise_fwd_bwd_stats(pkg, B, C, H, W, K, S, pad):
    from depth_b200 import ops
    x = rnd(B, C, H, W, seed=C + K)
    w = rnd(C, 1, K, K, seed=C + 7, scale=0.3)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, S, pad, 1, C)
    Ho, Wo = ref.shape[-2:]
    cot = rnd(B, C, Ho, Wo, seed=11)
    ref.backward(cot)
    xp = nhwc(x).requires_grad_(True)
    wp = w.cuda().requires_grad_(True)
    out, st = ops.dwconv(xp, wp, S, pad, pad, Ho, Wo, stats=True)
    ok, e, s = close(nchw(out), ref.detach())
    assert ok, (e, s)
    y = out.double().reshape(-1, C)
    tot = st.double().sum(0)
    assert torch.allclose(tot[0], y.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[1], (y * y).sum(0), rtol=1e-4, atol=1e-2)
    out.backward(nhwc(cot))
    ok, e, s = close(nchw(xp.grad), xr.grad)
    assert ok, ("dgrad", e, s)
    ok, e, s = close(wp.grad.cpu(), wr.grad, rel=2e-3, abs_frac=2e-3)
    assert ok, ("wgrad", e, s)


def test_depthwise_tf_same_padding(pkg):
    """asymmetric (TF 'SAME') geometry: pad_top/left = total // 2, the remainder falls on the bottom/right"""
    from depth_b200 import ops
    B, C, H, W, K, S = 1, 40, 22, 30, 3, 2
    x = rnd(B, C, H, W, seed=5)
    w = rnd(C, 1, K, K, seed=6, scale=0.3)
    Ho, Wo = -(-H // S), -(-W // S)
    ph, pw = max((Ho - 1) * S + K - H, 0), max((Wo - 1) * S + K - W, 0)
    ref = F.conv2d(F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)), w, None, S, 0, 1, C)
    out = ops.dwconv(nhwc(x), w.cuda(), S, ph // 2, pw // 2, Ho, Wo)
    ok, e, s = close(nchw(out), ref)
    assert ok, (e, s)


@pytest.mark.parametrize("B,C,H,W,K,S,pt,pl,pb,pr", [
    # stride-2 data / weight gradients for every parity of the top / left padding (the tap sets are compile-time per parity)
    (2, 72, 37, 46, 3, 2, 0, 0, 1, 1), (2, 72, 37, 46, 3, 2, 1, 0, 1, 1), (2, 72, 36, 45, 3, 2, 0, 1, 1, 0),
    (2, 72, 36, 45, 3, 2, 1, 1, 0, 0), (1, 136, 41, 70, 5, 2, 1, 2, 2, 1), (1, 136, 41, 70, 5, 2, 2, 1, 2, 2),
    (1, 136, 40, 71, 5, 2, 1, 1, 2, 2), (1, 136, 40, 71, 5, 2, 2, 2, 1, 1), (3, 144, 64, 96, 3, 2, 0, 0, 1, 1),
    # stride 1, asymmetric
    (2, 48, 30, 41, 5, 1, 1, 3, 3, 1), (2, 48, 30, 41, 3, 1, 0, 2, 2, 0)])
def test_depthwise_asymmetric_padding_fwd_bwd(pkg, B, C, H, W, K, S, pt, pl, pb, pr):
    from depth_b200 import ops
    x = rnd(B, C, H, W, seed=C + K + pt)
    w = rnd(C, 1, K, K, seed=C + 9, scale=0.3)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv2d(F.pad(xr, (pl, pr, pt, pb)), wr, None, S, 0, 1, C)
    Ho, Wo = ref.shape[-2:]
    cot = rnd(B, C, Ho, Wo, seed=13)
    ref.backward(cot)
    xp = nhwc(x).requires_grad_(True)
    wp = w.cuda().requires_grad_(True)
    out = ops.dwconv(xp, wp, S, pt, pl, Ho, Wo)
    ok, e, s = close(nchw(out), ref.detach())
    assert ok, (e, s)
    out.backward(nhwc(cot))
    ok, e, s = close(nchw(xp.grad), xr.grad)
    assert ok, ("dgrad", e, s)
    ok, e, s = close(wp.grad.cpu(), wr.grad, rel=2e-3, abs_frac=2e-3)
    assert ok, ("wgrad", e, s)


@pytest.mark.parametrize("C,res", [(32, False), (192, True), (1392, False)])
def test_batchnorm_relu6_fwd_bwd(pkg, C, res):
    from depth_b200 import ops
    B, H, W = 3, 12, 14
    x = rnd(B, C, H, W, seed=C, scale=3.0)
    r = rnd(B, C, H, W, seed=C + 1)
    bn = nn.BatchNorm2d(C, eps=1e-3)
    gp = torch.Generator().manual_seed(1000 + C)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=gp) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=gp) * 2 + 2)      # pushes a good share of the outputs above 6
    import copy
    bn_ref = copy.deepcopy(bn).train()
    xr = x.clone().requires_grad_(True)
    y_ref = bn_ref(xr)
    y_ref = F.relu6(y_ref + r) if res else F.relu6(y_ref)
    cot = rnd(B, C, H, W, seed=9)
    y_ref.backward(cot)
    bn_g = bn.cuda().train()
    xp = nhwc(x).requires_grad_(True)
    y = ops.bn_act(bn_g, xp, None, relu=2, res=nhwc(r) if res else None)
    ok, e, s = close(nchw(y), y_ref.detach())
    assert ok, (e, s)
    assert float((nchw(y) >= 6.0).float().mean()) > 0.02, "test must exercise the upper clamp"
    y.backward(nhwc(cot))
    ok, e, s = close(nchw(xp.grad), xr.grad, rel=2 ** -6, abs_frac=1e-2)
    assert ok, ("dx", e, s)
    ok, e, s = close(bn_g.weight.grad.cpu(), bn_ref.weight.grad, rel=1e-2, abs_frac=1e-2)
    assert ok, ("dgamma", e, s)
    ok, e, s = close(bn_g.bias.grad.cpu(), bn_ref.bias.grad, rel=1e-2, abs_frac=1e-2)
    assert ok, ("dbeta", e, s)
    assert torch.allclose(bn_g.running_var.cpu(), bn_ref.running_var, rtol=2e-2, atol=1e-3)


def test_stem_conv(pkg):
    from depth_b200 import ops
    B, H, W = 2, 36, 44
    x = rnd(B, 3, H, W, seed=2)
    w = rnd(32, 3, 3, 3, seed=3, scale=0.2)
    wr = w.clone().requires_grad_(True)
    ref = F.conv2d(x, wr, None, 2, 1)
    cot = rnd(*ref.shape, seed=4)
    ref.backward(cot)
    wp = w.cuda().requires_grad_(True)
    out, st = ops.stem_conv(x.cuda(), wp, stats=True)
    ok, e, s = close(nchw(out), ref.detach())
    assert ok, (e, s)
    out.backward(nhwc(cot))
    ok, e, s = close(wp.grad.cpu(), wr.grad, rel=2e-3, abs_frac=2e-3)
    assert ok, ("wgrad", e, s)


def _trunk_pair():
    import copy
    from depth_b200 import standins
    from depth_b200.network import blocks, encoder_fused
    blocks.hub_load = standins.hub_load_standin
    torch.manual_seed(0)
    ref = blocks._make_pretrained_efficientnet_lite3(False)
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 4:
                p.copy_(p.to(BF).float())
    fused = copy.deepcopy(ref).cuda().train()
    return ref.cuda().train(), fused, encoder_fused


def test_trunk_blocks_match_pytorch(pkg):
    """every block of the EfficientNet-Lite3-shaped trunk in train mode, fed the SAME bf16 input on both sides: fused
    kernels vs PyTorch fp32.  Per block: output within 2 % relative L2 (three bf16-stored intermediate tensors +
    train-mode BN); input gradient within 8 % relative L2 - the ReLU6 gates are decided on bf16-stored pre-activations,
    so the ~0.4 % of elements within rounding distance of 0 or 6 gate differently from the fp32 run, each contributing
    its full gradient (sqrt(0.004) ~ 6 %); parameter gradients direction-faithful (cosine > 0.98); BN running
    statistics within 1 %."""
    ref, fused, ef = _trunk_pair()
    assert ef.supported(fused)
    rb = ef._walk(ref.layer1)[3:] + ef._walk(ref.layer2) + ef._walk(ref.layer3) + ef._walk(ref.layer4)
    fb = ef._walk(fused.layer1)[3:] + ef._walk(fused.layer2) + ef._walk(fused.layer3) + ef._walk(fused.layer4)
    assert len(rb) == len(fb) == 24
    H, W = 48, 64
    for i, (r, f) in enumerate(zip(rb, fb)):
        cin = (r.conv_pw if hasattr(r, "conv_pwl") else r.conv_dw).in_channels
        stride = r.conv_dw.stride[0]
        x = rnd(2, cin, H, W, seed=100 + i).cuda()
        xr = x.clone().requires_grad_(True)
        yr = r(xr)
        xf = x.permute(0, 2, 3, 1).contiguous().to(BF).requires_grad_(True)
        yf = ef.run_block(f, xf)
        got = yf.float().permute(0, 3, 1, 2)
        rel = float((got - yr).norm() / yr.norm())
        assert rel < 0.02, (i, "fwd", rel)
        cot = rnd(*yr.shape, seed=200 + i).cuda()
        yr.backward(cot)
        yf.backward(cot.permute(0, 2, 3, 1).contiguous().to(BF))
        rel = float((xf.grad.float().permute(0, 3, 1, 2) - xr.grad).norm() / xr.grad.norm())
        assert rel < 0.08, (i, "dx", rel)
        # parameter gradients: self-calibrated against stock bf16 autocast of the same block on the same input - the
        # fused path may not be further from the fp32 gradients than twice the autocast run's own distance (+ 0.01)
        import copy
        ra = copy.deepcopy(r)
        for p_ in ra.parameters():
            p_.grad = None
        xa = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=BF):
            ya = ra(xa)
        ya.backward(cot.to(ya.dtype))
        for (n, pf), (_, pr), (_, pa) in zip(f.named_parameters(), r.named_parameters(), ra.named_parameters()):
            cos = float(F.cosine_similarity(pf.grad.flatten().double(), pr.grad.flatten().double(), dim=0))
            cos_a = float(F.cosine_similarity(pa.grad.flatten().double(), pr.grad.flatten().double(), dim=0))
            assert 1.0 - cos <= 2.0 * (1.0 - cos_a) + 0.01, (i, n, cos, cos_a)
        for (n, bf_), (_, br) in zip(f.named_buffers(), r.named_buffers()):
            if n.endswith("running_var") or n.endswith("running_mean"):
                assert torch.allclose(bf_, br, rtol=1e-2, atol=1e-2), (i, n)
        if stride == 2:
            H, W = max(H // 2, 12), max(W // 2, 16)


def test_trunk_end_to_end(pkg):
    """whole trunk, train mode, 4 x 256x320 input: bf16 storage noise is amplified by ~75 train-mode BatchNorms over few
    samples per channel in the deep stages, so the end-to-end bound is calibrated against stock bf16 autocast of the
    same module; the per-block test above carries the tight tolerance."""
    ref, fused, ef = _trunk_pair()
    x = rnd(4, 3, 256, 320, seed=1).cuda()
    feats = ef.forward(fused, x)
    r1 = ref.layer1(x); r2 = ref.layer2(r1); r3 = ref.layer3(r2); r4 = ref.layer4(r3)
    import copy
    auto = copy.deepcopy(ref)
    with torch.autocast("cuda", dtype=BF):
        a1 = auto.layer1(x); a2 = auto.layer2(a1); a3 = auto.layer3(a2); a4 = auto.layer4(a3)
    for f, r, a in zip(feats, [r1, r2, r3, r4], [a1, a2, a3, a4]):
        assert tuple(f.shape) == (r.shape[0], r.shape[2], r.shape[3], r.shape[1])
        rel = float((f.float().permute(0, 3, 1, 2) - r).norm() / r.norm())
        rel_auto = float((a.float() - r).norm() / r.norm())
        # self-calibrated: no further from fp32 than 1.5x stock bf16 autocast's own drift (+ 2 %)
        assert rel <= 1.5 * rel_auto + 0.02, (rel, rel_auto)
    sum(f.float().sum() for f in feats).backward()
    for n, p in fused.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
