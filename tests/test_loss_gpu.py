"""GPU parity of the fused loss / metric kernels (through the C ABI) against the oracle and the
golden values computed by the reference's own util.py.  Tolerance: 1e-5 relative in fp32 (north star);
delta pixel counts within 0.01 % of pixels."""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures as fx, losses as ol
from oracle.make_golden import loss_cases, THRESH

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL = 1e-5


def _close(a, b, rel=REL, abs_=1e-7):
    a, b = float(a), float(b)
    if np.isnan(b):
        return np.isnan(a)
    return abs(a - b) <= rel * abs(b) + abs_


@pytest.mark.parametrize("name", ["survey", "bench_like", "small_zeros", "tiny"])
def test_forward_values_vs_reference_golden(pkg, golden_loss, name):
    p, t, rgb = loss_cases()[name]()
    pc, tc, rc = p.cuda(), t.cuda(), rgb.cuda()
    exp = golden_loss[name]["fp32"]
    exp64 = golden_loss[name]["fp64"]
    got = {
        "si": pkg.scale_invariant_loss(pc, tc).item(),
        "si_sqrt": pkg.scale_invariant_loss(pc, tc, sqroot=True).item(),
        "silog": pkg.silog_loss(pc, tc, mask=(tc > 0)).item(),
        "grad": pkg.gradient_loss(pc, tc).item(),
        "edge": pkg.edge_aware_loss(pc, tc, rc, 0.5).item(),
        "absrel": pkg.absolute_relative_error(pc, tc).item(),
    }
    for i, th in enumerate(THRESH):
        got[f"delta{i}"] = pkg.delta_thres(pc, tc, thres=th).item()
    for k in exp:
        # fp32 reference value, or the fp64 evaluation of the same formula: we must be within 1e-5 of the reference
        assert _close(got[k], exp[k]) or _close(got[k], exp64[k]), (name, k, got[k], exp[k], exp64[k])
    # integer counts: within 0.01 % of pixels per sample
    cnt = pkg.delta_counts(pc, tc, THRESH).cpu().numpy()
    ref = np.array(golden_loss[name]["delta_counts"])
    npx = p[0].numel()
    assert np.abs(cnt - ref).max() <= max(1, int(1e-4 * npx)), (cnt, ref)


@pytest.mark.parametrize("name", ["small_zeros", "tiny"])
def test_gradients_vs_reference_stored(pkg, name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "loss_small_inputs.npz"))
    p, t, rgb = (torch.from_numpy(z[f"{name}.{k}"]).cuda() for k in ("pred", "target", "rgb"))
    fns = {"si": lambda q: pkg.scale_invariant_loss(q, t), "silog": lambda q: pkg.silog_loss(q, t, mask=(t > 0)),
           "grad": lambda q: pkg.gradient_loss(q, t), "edge": lambda q: pkg.edge_aware_loss(q, t, rgb, 0.5)}
    for k, fn in fns.items():
        q = p.clone().requires_grad_(True)
        fn(q).backward()
        ref = torch.from_numpy(z[f"{name}.grad_{k}"]).cuda()
        scale = float(ref.abs().max())
        assert float((q.grad - ref).abs().max()) <= 2e-5 * scale + 1e-9, (name, k)


def test_gradients_vs_oracle_full_size(pkg):
    p, t, rgb = loss_cases()["bench_like"]()
    cfg = fx.loss_config(si=1.0, silog=1.0, vf=0.85, grad=0.2, edge=0.5)
    q = p.clone().requires_grad_(True)
    tot, d = ol.combined_loss(q, t, cfg, rgb=rgb)
    tot.backward()
    qc = p.cuda().requires_grad_(True)
    totc, dc = pkg.combined_loss(qc, t.cuda(), cfg, rgb=rgb.cuda())
    (totc * 1.0).backward()
    assert _close(totc.item(), tot.item())
    for k in d:
        assert _close(dc[k], d[k]), (k, dc[k], d[k])
    ref = q.grad
    err = (qc.grad.cpu() - ref).abs()
    # sign() terms of the stencil losses flip only where |a|-|b| is within rounding of 0: allow a vanishing fraction
    tol = 2e-5 * ref.abs() + 1e-5 * float(ref.abs().mean())
    assert float((err > tol).float().mean()) < 1e-5
    cfg0 = fx.loss_config()           # config.yaml defaults 1/0/0/0
    totc0, dc0 = pkg.combined_loss(p.cuda(), t.cuda(), cfg0, rgb=rgb.cuda())
    tot0, d0 = ol.combined_loss(p, t, cfg0, rgb=rgb)
    assert _close(totc0.item(), tot0.item()) and dc0["silog_loss"] == 0.0 and dc0["edge_loss"] == 0.0


def test_evaluation_metrics_fused(pkg):
    """the fused evaluation.py metric set equals the five separate reference calls."""
    p, t, _ = loss_cases()["bench_like"]()
    out = pkg.evaluation_metrics(p.cuda(), t.cuda(), thresholds=[1.05, 1.05 ** 2, 1.05 ** 3]).cpu()
    exp = [ol.scale_invariant_loss(p, t, sqroot=True), ol.absolute_relative_error(p, t)] + \
          [ol.delta_thres(p, t, 1.05 ** j) for j in (1, 2, 3)]
    for a, b in zip(out.tolist(), exp):
        assert _close(a, b.item(), rel=1e-5, abs_=2e-6)


def test_evaluate_model_metric_set(pkg, golden_loss):
    """main.evaluate_model sums (MAE/RMSE/REL/siRMSE/unaligned delta<1.25^k)."""
    p, t, _ = loss_cases()["small_zeros"]()
    got = pkg.evaluate_model_sums(p.cuda(), t.cuda())
    exp = ol.evaluate_metric_sums(p, t)
    for k in ("abs", "sq", "rel", "sirmse"):
        assert _close(got[k], exp[k], rel=2e-5), (k, got[k], exp[k])
    for k in ("d1", "d2", "d3"):
        assert abs(got[k] - exp[k]) <= 1


def test_properties_full_size(pkg):
    """size-independent properties at bench size (B=32, 448x576): scale invariance of SI and of aligned delta,
    SI(p,p)=0, delta(p,p)=1, sample-permutation invariance."""
    g = torch.Generator(device="cuda").manual_seed(7)
    B, H, W = 32, 448, 576
    t = torch.rand(B, 1, H, W, device="cuda", generator=g) * 9.9 + 0.1
    p = t * torch.exp(0.2 * torch.randn(B, 1, H, W, device="cuda", generator=g))
    si = pkg.scale_invariant_loss(p, t).item()
    si_scaled = pkg.scale_invariant_loss(p * 3.7, t).item()
    assert abs(si - si_scaled) <= 2e-4 * abs(si)          # eps=1e-6 breaks exact invariance only slightly
    assert abs(pkg.scale_invariant_loss(p, p).item()) < 1e-7
    assert pkg.delta_thres(p, p, thres=1.05).item() == 1.0
    d = pkg.delta_thres(p, t, thres=1.25).item()
    d_scaled = pkg.delta_thres(p * 0.31, t, thres=1.25).item()
    assert abs(d - d_scaled) < 1e-4
    perm = torch.randperm(B, device="cuda")
    assert abs(pkg.scale_invariant_loss(p[perm], t[perm]).item() - si) <= 1e-6 * abs(si)
    # analytic check: d ~ N(0, 0.2^2) so SI ~ 0.04
    assert abs(si - 0.04) < 1e-3


def test_shape_assert(pkg):
    a = torch.rand(2, 1, 8, 8, device="cuda")
    b = torch.rand(2, 1, 8, 9, device="cuda")
    with pytest.raises(AssertionError):
        pkg.scale_invariant_loss(a, b)
    with pytest.raises(AssertionError):
        pkg.delta_thres(a, b)


@pytest.mark.parametrize("B,H,W", [(3, 37, 53), (2, 64, 96), (9, 448, 576)])
def test_fused_eval_counts_match_two_pass_counts(pkg, B, H, W):
    """the fused kernels (37x53: pixel count not a multiple of 4 -> the thread-block-cluster kernel with its DSMEM
    exchange of the per-sample scale; the others -> the streaming kernel) must classify exactly the pixels the separate
    moments + delta_counts passes classify: integer equality of the per-batch counts."""
    g = torch.Generator().manual_seed(B * H + W)
    t = (torch.rand(B, 1, H, W, generator=g) * 9.9 + 0.1).cuda()
    p = (t.cpu() * torch.exp(0.1 * torch.randn(B, 1, H, W, generator=g)) * 1.3).cuda()
    p[:, :, 3:9, 5:17] = 0.0
    thr = [1.05, 1.05 ** 2, 1.05 ** 3]
    out = pkg.evaluation_metrics(p, t, thresholds=thr, fast_math=False).double().cpu()
    cnt = pkg.delta_counts(p, t, thr, aligned=True).cpu()
    n = H * W
    want = (cnt.double() / n).float().double().mean(0)
    assert torch.allclose(out[2:], want, rtol=0, atol=1e-7), (out[2:], want)
    assert abs(out[0].item() - pkg.scale_invariant_loss(p, t, sqroot=True).item()) <= 1e-6
    assert abs(out[1].item() - pkg.absolute_relative_error(p, t).item()) <= 1e-6 * out[1].item()


@pytest.mark.parametrize("B,H,W,thr", [
    (40, 448, 576, [1.05, 1.05 ** 2, 1.05 ** 3]),      # several samples per CTA group: mbarrier phases, chunk refill
    (5, 426, 560, [1.25]),                             # raw test-set resolution, ragged last slice, one threshold
    (3, 448, 576, [1.1, 1.3]),                         # run-time threshold count, one-division form
    (20, 448, 576, [1.1, 1.3, 1.5, 2.0, 3.0]),         # run-time count with full-size slices (largest static smem)
    (3, 128, 160, [0.9, 1.0, 1.2, 1.5]),               # thresholds <= 1: both quotients as written in util.py:204
    (2, 896, 1152, [1.05, 1.05 ** 2, 1.05 ** 3]),      # config-5 resolution: more CTAs per sample
])
def test_streaming_eval_counts_exact(pkg, B, H, W, thr):
    """the streaming kernel (slices in shared memory, one IEEE division per pixel when every threshold is > 1) must
    classify exactly the pixels the two-pass path (both quotients, util.py:204-205) classifies - integer equality -
    including zero predictions, zero targets and fast_math staying inside the contract."""
    g = torch.Generator().manual_seed(B * H + W)
    t = (torch.rand(B, 1, H, W, generator=g) * 9.9 + 0.1)
    p = t * torch.exp(0.1 * torch.randn(B, 1, H, W, generator=g)) * 1.3
    p[:, :, 3:9, 5:17] = 0.0
    t[:, :, 20:23, 40:47] = 0.0
    p[0, :, 21, 41] = 0.0           # 0 / 0
    p, t = p.cuda(), t.cuda()
    out = pkg.evaluation_metrics(p, t, thresholds=thr, fast_math=False).double().cpu()
    cnt = pkg.delta_counts(p, t, thr, aligned=True).cpu()
    n = H * W
    want = (cnt.double() / n).float().double().mean(0)
    assert torch.allclose(out[2:], want, rtol=0, atol=1e-7), (out[2:], want)
    assert abs(out[0].item() - pkg.scale_invariant_loss(p, t, sqroot=True).item()) <= 1e-6
    assert abs(out[1].item() - pkg.absolute_relative_error(p, t).item()) <= 1e-6 * abs(out[1].item())
    fast = pkg.evaluation_metrics(p, t, thresholds=thr, fast_math=True).double().cpu()
    assert abs(fast[0] - out[0]) <= 1e-5 * abs(out[0])
    assert float((fast[2:] - out[2:]).abs().max()) <= 1e-4
    # same call again: the arrival counters are reset by the entry point, results are deterministic
    again = pkg.evaluation_metrics(p, t, thresholds=thr, fast_math=False).double().cpu()
    assert torch.equal(again, out)
    # the default (lean) arithmetic on the same inputs, zeros and 0/0 included: inside the contract
    lean = pkg.evaluation_metrics(p, t, thresholds=thr).double().cpu()
    assert abs(lean[0] - out[0]) <= 1e-5 * abs(out[0]) and abs(lean[1] - out[1]) <= 1e-5 * abs(out[1])
    assert float((lean[2:] - out[2:]).abs().max()) <= 1e-4


def test_fused_eval_fast_math_within_contract(pkg):
    """MUFU variant of the fused evaluation kernel: SI-RMSE / AbsRel within 1e-5 relative of the exact path and the
    delta fractions within 0.01 % of pixels (the north-star tolerances; measured differences are ~1e-7 / a few ppm)."""
    B, H, W = 8, 448, 576
    g = torch.Generator().manual_seed(77)
    t = (torch.rand(B, 1, H, W, generator=g) * 9.9 + 0.1).cuda()
    p = (t.cpu() * torch.exp(0.1 * torch.randn(B, 1, H, W, generator=g)) * 1.3).cuda()
    p[:, :, 100:140, 200:300] = 0.0
    thr = [1.05, 1.05 ** 2, 1.05 ** 3]
    exact = pkg.evaluation_metrics(p, t, thresholds=thr, fast_math=False).double().cpu()
    fast = pkg.evaluation_metrics(p, t, thresholds=thr, fast_math=True).double().cpu()
    assert abs(fast[0] - exact[0]) <= 1e-5 * abs(exact[0])
    assert abs(fast[1] - exact[1]) <= 1e-5 * abs(exact[1])
    assert float((fast[2:] - exact[2:]).abs().max()) <= 1e-4


def test_evaluation_loop_matches_reference_arithmetic(pkg):
    """depth_b200.evaluation.evaluate_batches vs the reference's loop (evaluation.py:138-186) re-enacted with the oracle
    functions: batch-length weighting and the N_SAMPLES clip of the last batch."""
    import torch.nn as nn

    class Identity(nn.Module):
        def forward(self, x):          # "model": the rgb tensor already is the prediction (B,1,H,W)
            return x

    g = torch.Generator().manual_seed(5)
    batches = []
    for b in (4, 4, 3):
        t = torch.rand(b, 1, 48, 64, generator=g) * 9.9 + 0.1
        p = t * torch.exp(0.1 * torch.randn(b, 1, 48, 64, generator=g)) * 1.2
        batches.append((p, t, None))
    n_samples = 10
    got = pkg.evaluation.evaluate_batches(Identity(), batches, torch.device("cuda"), n_samples=n_samples)
    tot = [0.0] * 5
    seen = 0
    for p, t, _ in batches:
        b = p.shape[0]
        take = max(0, min(b, n_samples - seen)); seen += b
        vals = [ol.scale_invariant_loss(p, t, sqroot=True).item(), ol.absolute_relative_error(p, t).item()] + \
               [ol.delta_thres(p, t, 1.05 ** j).item() for j in (1, 2, 3)]
        tot = [a + v * take for a, v in zip(tot, vals)]
    want = [v / n_samples for v in tot]
    assert got["samples"] == n_samples
    assert abs(got["si_rmse"] - want[0]) <= 1e-5 * abs(want[0])
    assert abs(got["abs_rel"] - want[1]) <= 1e-5 * abs(want[1])
    assert all(abs(a - b) <= 1e-4 for a, b in zip(got["delta"], want[2:]))


def test_evaluate_model_matches_reference_golden(pkg, golden_loss):
    """M4 end to end: depth_b200.evaluation.evaluate_model (fused moments + counting kernels) against the dict produced by
    the reference's own main.evaluate_model on the stored inputs.  Scalars within 1e-5 relative; the delta fractions
    within 0.01 % of pixels."""
    import os
    import numpy as np
    store = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_small_inputs.npz"))

    class Ident(torch.nn.Module):
        def forward(self, x):
            return x[:, 0]

    batches = [(torch.from_numpy(store[f"evaluate_model.inputs{i}"]), torch.from_numpy(store[f"evaluate_model.targets{i}"]), None)
               for i in range(2)]
    got = pkg.evaluation.evaluate_model(Ident(), batches, torch.device("cuda"))
    ref = golden_loss["evaluate_model"]
    for k in ("MAE", "RMSE", "siRMSE", "REL"):
        assert abs(got[k] - ref[k]) <= 1e-5 * abs(ref[k]), (k, got[k], ref[k])
    for k in ("Delta1", "Delta2", "Delta3"):
        assert abs(got[k] - ref[k]) <= 1e-4, (k, got[k], ref[k])


@pytest.mark.parametrize("case", ["positive", "zeros", "negative", "nan", "thr_le_1"])
def test_default_eval_arithmetic_within_contract(pkg, case):
    """The default (lean) arithmetic of evaluation_metrics - one shared reciprocal + one lg2 per pixel, division-free
    threshold test - against the exact IEEE path (the reference's arithmetic) at the evaluation size: SI-RMSE / AbsRel
    within 1e-5 relative, delta fractions within 0.01 % of pixels (the north-star contract), for ordinary depth maps, exact
    zeros in either operand (x/0, 0/0), negative values (the slice falls back to the two-quotient code: sign semantics of
    util.py:204-205), NaNs (the sample's scale is NaN: every count 0, SI-RMSE NaN, as in the reference) and thresholds
    <= 1 (never reached)."""
    B, H, W = 6, 448, 576
    g = torch.Generator().manual_seed(321)
    t = torch.rand(B, 1, H, W, generator=g) * 9.9 + 0.1
    p = t * torch.exp(0.1 * torch.randn(B, 1, H, W, generator=g)) * 1.3
    thr = [1.05, 1.05 ** 2, 1.05 ** 3]
    if case == "zeros":
        p[:, :, 50:90, 100:300] = 0.0
        t[:, :, 200:230, 40:470] = 0.0
        p[:, :, 210:215, 50:60] = 0.0
    elif case == "negative":
        p[1, :, 17, 33] = -0.5          # log of a negative number: that sample's metrics are NaN in the reference too
        t[3, :, 100:110, 200:260] = -1e-7   # inside (-eps, 0): finite logs, negative quotients
    elif case == "nan":
        p[2, :, 5, 5] = float("nan")
    elif case == "thr_le_1":
        thr = [0.9, 1.0, 1.2]
    p, t = p.cuda(), t.cuda()
    exact = pkg.evaluation_metrics(p, t, thresholds=thr, fast_math=False).double().cpu()
    lean = pkg.evaluation_metrics(p, t, thresholds=thr).double().cpu()
    for i in range(2):
        if torch.isnan(exact[i]):
            assert torch.isnan(lean[i]), (case, i, lean, exact)
        else:
            assert abs(lean[i] - exact[i]) <= 1e-5 * abs(exact[i]), (case, i, lean, exact)
    assert float((lean[2:] - exact[2:]).abs().max()) <= 1e-4, (case, lean, exact)


def test_per_pixel_scale_invariant_loss(pkg):
    """M5 (util.py:159-181): the per-pixel map against the oracle's restatement (pinned to the reference in
    oracle/make_golden.py's loss cases), single image as the reference's visualiser calls it, 1e-5 of the map's maximum."""
    g = torch.Generator().manual_seed(11)
    t = torch.rand(448, 576, generator=g) * 9.9 + 0.1
    p = t * torch.exp(0.2 * torch.randn(448, 576, generator=g)) * 1.3
    got = pkg.per_pixel_scale_invariant_loss(p.cuda(), t.cuda()).cpu()
    ref = ol.per_pixel_scale_invariant_loss(p, t)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.max())
    with pytest.raises(AssertionError):
        pkg.per_pixel_scale_invariant_loss(-p.cuda(), t.cuda())
    # the map the REFERENCE's own function produced (tests/golden, written by oracle/make_golden.py)
    import os
    import numpy as np
    store = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_small_inputs.npz"))
    got = pkg.per_pixel_scale_invariant_loss(torch.from_numpy(store["per_pixel_si.pred"]).cuda(),
                                             torch.from_numpy(store["per_pixel_si.target"]).cuda()).cpu()
    ref = torch.from_numpy(store["per_pixel_si.map"])
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.max())
