"""Parity at the benched configuration (default MidasNetSemantics, 448x576, bf16) and over several optimisation steps -
the path bench.py measures (fused EfficientNet trunk, CUDA-graph replay + AdamW), not a tiny fixture.

Oracle: oracle/model.py (pinned to the reference by oracle/make_golden.py) run by PyTorch in fp32 on the same GPU with
TF32 disabled.

What can be stated absolutely and what cannot (measured, tools/diag_fullres.py): with the third-party trunk kept in fp32
the whole in-scope path (reassemble convs, four fusion blocks, DINOv2 head, cross attention, ResidualBlocks, depth head -
about forty bf16 layers) reproduces the fp32 oracle to 1.5e-2 max-norm / 6e-3 relative L2 at 448x576.  The RANDOM-INIT
EfficientNet-Lite3 stand-in is chaotic in bf16: its stride-16 / stride-32 maps drift 19 % / 55 % in relative L2 from fp32
for ANY bf16 execution - PyTorch's own bf16 autocast of the oracle drifts 0.180 at the output, our fused trunk 0.179 -
so with the fused trunk the yardstick is relative (no worse than stock autocast) and the trunk kernels are pinned layer
by layer in tests/test_encoder_gpu.py.  Tolerances:

  trunk in fp32:  forward depth map max-norm <= 3e-2 of the map's maximum, relative L2 <= 1.5e-2;
                  SI loss within 5e-3 relative;
                  eval-mode BN, positive linear functional: every in-scope parameter gradient <= 5e-2 relative L2,
                  median <= 1e-2;
                  train-mode BN + SI loss: refinenets / output_conv / ResidualBlocks / heads <= 0.10 each, median <= 5e-2;
                  the cross-attention / DINOv2-head / reassemble / trunk gradients are ill-conditioned in fp32 itself
                  (stock bf16 autocast drifts > 100 % there): ours <= 0.5 x stock autocast's drift per tensor;
                  BN running statistics <= 2e-3 relative after one step, num_batches_tracked identical
  fused trunk:    output drift (relative L2) <= 1.1 x the drift of the oracle under torch.autocast(bf16); loss within 2e-2
  5 optimisation steps:  graph replay == eager step (identical loss scalars, bit-identical parameters and buffers);
                  loss trajectory within 2e-2 of the fp32 oracle's AdamW trajectory; with the fp32 trunk also BN
                  running_* after the first step within 5e-3 (trunk layers 3e-2) and num_batches_tracked identical (the shared spatial_reduction BN counts 2 / step)
"""
import copy

import numpy as np
import pytest
import torch

from oracle import cases, fixtures as fx, losses as ol

pytestmark = pytest.mark.gpu

H, W = 448, 576


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _pair(pkg):
    from depth_b200 import standins
    from depth_b200.network import blocks, midas_semantics
    blocks.hub_load = standins.hub_load_standin
    ora = cases.build_oracle_semantics(standins)
    cases.prepare_full(ora)
    with torch.no_grad():
        for p in ora.parameters():
            if p.dim() == 4:
                p.copy_(p.to(torch.bfloat16).float())
    prod = midas_semantics.MidasNetSemantics(None, features=64, backbone="efficientnet_lite3", exportable=True,
                                             non_negative=True, cfg=fx.model_cfg(), blocks={'expand': True},
                                             dinov2_type='dinov2_vits14')
    prod.load_state_dict(ora.state_dict(), strict=True)
    assert prod.fused_encoder, "the default (and benched) path runs the trunk on the sm_100a kernels"
    return ora.cuda().train(), prod.cuda().train()


def _batch(B, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, H, W, generator=g).to(torch.bfloat16).float()
    t = torch.rand(B, 1, H, W, generator=g) * 9.9 + 0.1
    return x.cuda(), t.cuda()


def rel_l2(a, b):
    return float((a.double() - b.double()).norm()) / max(float(b.double().norm()), 1e-30)


def rel_max(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def test_forward_backward_at_448x576(pkg):
    ora, prod = _pair(pkg)
    prod.fused_encoder = False          # third-party trunk in fp32 on both sides: absolute tolerances (module docstring)
    auto = copy.deepcopy(ora)
    x, t = _batch(4)
    out_o = ora(x)
    loss_o = ol.scale_invariant_loss(out_o.unsqueeze(1), t)
    loss_o.backward()
    out_p = prod(x)
    loss_p = pkg.scale_invariant_loss(out_p.unsqueeze(1), t)
    loss_p.backward()
    torch.cuda.synchronize()
    e_max, e_l2 = rel_max(out_p.detach(), out_o.detach()), rel_l2(out_p.detach(), out_o.detach())
    print(f"forward: max-norm {e_max:.4f}, rel L2 {e_l2:.4f}; loss {loss_p.item():.6f} vs {loss_o.item():.6f}")
    assert e_max <= 3e-2 and e_l2 <= 1.5e-2
    assert abs(loss_p.item() - loss_o.item()) <= 5e-3 * abs(loss_o.item())
    # the same step under PyTorch's stock bf16 autocast: the yardstick for the ill-conditioned part (below)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out_a = auto(x)
    ol.scale_invariant_loss(out_a.float().unsqueeze(1), t).backward()
    go, ga = dict(ora.named_parameters()), dict(auto.named_parameters())
    errs, stock = {}, {}
    gmax = max(float(p.grad.norm()) for p in go.values() if p.grad is not None)
    for k, p in prod.named_parameters():
        if k.startswith("dinov2."):
            continue
        if go[k].grad is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        n = float(go[k].grad.norm())
        if n < 1e-6 * gmax:          # exactly-zero true gradients (conv bias in front of a train-mode BN)
            continue
        errs[k] = rel_l2(p.grad, go[k].grad)
        stock[k] = rel_l2(ga[k].grad.float(), go[k].grad)
    # (a) layers whose gradient does not pass through the cross-attention stack (six train-mode BatchNorms around a
    #     softmax, fed by the scale-invariant loss' zero-mean gradient): absolute tolerance
    direct = ("scratch.refinenet", "scratch.output_conv", "fusion_blocks.", "fusion_head.", "depth_head.")
    d = {k: e for k, e in errs.items() if k.startswith(direct)}
    worst = sorted(d.items(), key=lambda kv: -kv[1])[:3]
    print(f"train-mode SI gradients, direct layers: {len(d)} tensors, median {np.median(list(d.values())):.4f}, worst {worst}")
    assert worst[0][1] <= 0.10 and float(np.median(list(d.values()))) <= 5e-2, worst
    # (b) everything else (DINOv2 head, cross attention, reassemble convs, the trunk): the fp32 problem itself is
    #     ill-conditioned there - stock bf16 autocast is off by > 100 % - so the statement is relative: at most half of
    #     stock autocast's drift on the same tensor (measured: 3x - 10x better)
    rest = {k: (e, stock[k]) for k, e in errs.items() if not k.startswith(direct)}
    bad = [(k, e, sa) for k, (e, sa) in rest.items() if e > max(0.5 * sa, 2e-2)]
    ratio = float(np.median([e / max(sa, 1e-9) for e, sa in rest.values()]))
    print(f"train-mode SI gradients, remaining {len(rest)} tensors: median drift ratio ours / stock autocast {ratio:.3f}")
    assert not bad, bad[:5]
    bo = dict(ora.named_buffers())
    for k, b in prod.named_buffers():
        if k.startswith("dinov2."):
            continue
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(bo[k]), k
        else:
            assert rel_max(b.float(), bo[k].float()) <= 2e-3, (k, rel_max(b.float(), bo[k].float()))


def test_eval_mode_gradients_absolute(pkg):
    """The well-conditioned gradient check at the benched resolution: eval-mode BatchNorm (no batch-mean cancellation) and
    a positive linear functional of the depth map.  Every in-scope parameter tensor within 5e-2 relative L2 of the fp32
    oracle (measured worst 3.3e-2 in the DINOv2 head, most layers 1e-3), median <= 1e-2."""
    ora, prod = _pair(pkg)
    prod.fused_encoder = False
    ora.eval(); prod.eval()
    x, _ = _batch(4)
    w = (torch.rand(4, H, W, generator=torch.Generator().manual_seed(5)) + 0.5).cuda()
    (ora(x) * w).mean().backward()
    (prod(x) * w).mean().backward()
    go = dict(ora.named_parameters())
    gmax = max(float(p.grad.norm()) for p in go.values() if p.grad is not None)
    errs = {}
    for k, p in prod.named_parameters():
        if k.startswith(("dinov2.", "pretrained.")) or go[k].grad is None:
            continue
        if float(go[k].grad.norm()) < 1e-6 * gmax:
            continue
        errs[k] = rel_l2(p.grad, go[k].grad)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    med = float(np.median(list(errs.values())))
    print(f"eval-mode gradients: {len(errs)} tensors, median {med:.4f}, worst {worst}")
    assert worst[0][1] <= 5e-2 and med <= 1e-2, worst


def test_fused_trunk_drift_not_worse_than_stock_autocast(pkg):
    ora, prod = _pair(pkg)
    x, t = _batch(4)
    with torch.no_grad():
        ref = ora(x)
        out = prod(x)                                  # default path: fused trunk
        with torch.autocast("cuda", dtype=torch.bfloat16):
            auto = ora(x).float()
    d_ours, d_auto = rel_l2(out, ref), rel_l2(auto, ref)
    l_ref = ol.scale_invariant_loss(ref.unsqueeze(1), t).item()
    l_ours = pkg.scale_invariant_loss(out.unsqueeze(1), t).item()
    print(f"output drift from fp32: ours {d_ours:.4f}, stock bf16 autocast {d_auto:.4f}; loss {l_ours:.5f} vs {l_ref:.5f}")
    assert d_ours <= 1.1 * d_auto
    assert abs(l_ours - l_ref) <= 2e-2 * abs(l_ref)


def _opt(model):
    return torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, fused=True,
                             capturable=True)


@pytest.mark.parametrize("fused", [True, False])
def test_five_steps_graph_vs_eager_vs_oracle(pkg, fused):
    """reference loop main.py:125-144: zero_grad, forward, combined_loss, backward, AdamW step - five times."""
    ora, prod = _pair(pkg)
    prod.fused_encoder = fused
    eager = copy.deepcopy(prod)
    cfg = fx.loss_config()
    B, steps = 2, 5
    batches = [_batch(B, seed=100 + i) for i in range(steps)]
    # --- oracle (fp32) ---
    opt_o = torch.optim.AdamW([p for p in ora.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
    loss_o = []
    bo1 = None
    for x, t in batches:
        opt_o.zero_grad(set_to_none=True)
        l = ol.scale_invariant_loss(ora(x).unsqueeze(1), t)
        l.backward()
        opt_o.step()
        loss_o.append(l.item())
        if bo1 is None:
            bo1 = {k: v.detach().clone() for k, v in ora.named_buffers()}
    # --- eager product steps (same body the graph captures, dispatched kernel by kernel) ---
    opt_e = _opt(eager)
    loss_e = []
    for x, t in batches:
        for p in eager.parameters():
            p.grad = None
        total, out = pkg.util.combined_loss_device(eager(x).unsqueeze(1), t, cfg, rgb=x)
        total.backward()
        opt_e.step()
        loss_e.append(out[pkg._lib.L_TOTAL].item())
    # --- graph replays ---
    opt_g = _opt(prod)
    before = {k: v.detach().clone() for k, v in prod.state_dict().items()}
    gstep = pkg.GraphedTrainStep(prod, opt_g, cfg, batches[0][0], batches[0][1], use_rgb=True, world=1, warmup=2)
    for k, v in prod.state_dict().items():        # the warm-up steps must leave no trace
        assert torch.equal(v, before[k]), f"GraphedTrainStep construction changed {k}"
    loss_g = []
    bg1 = None
    for x, t in batches:
        gstep(x, t)
        loss_g.append(gstep.loss_dict()["total"])
        if bg1 is None:
            torch.cuda.synchronize()
            bg1 = {k: v.detach().clone() for k, v in prod.named_buffers()}
    torch.cuda.synchronize()
    print("loss oracle", loss_o, "\nloss eager ", loss_e, "\nloss graph ", loss_g)
    se, sg = eager.state_dict(), prod.state_dict()
    if fused:
        assert loss_g == loss_e, "graph replay must reproduce the eager step exactly"
        for k in se:
            assert torch.equal(se[k], sg[k]), f"graph vs eager: {k} differs after {steps} steps"
    else:
        # the PyTorch-run trunk is not bit-reproducible between eager dispatch and graph capture (cuDNN picks its
        # algorithms per call): and the random-init network amplifies the
        # last-bit differences step by step: the two trajectories agree to 5e-3 instead of bit for bit
        for a, b in zip(loss_g, loss_e):
            assert abs(a - b) <= 5e-3 * abs(b), (loss_g, loss_e)
    for a, b in zip(loss_g, loss_o):
        assert abs(a - b) <= 2e-2 * abs(b), (loss_g, loss_o)
    bo = dict(ora.named_buffers())
    for k, b in prod.named_buffers():
        if k.startswith("dinov2."):
            continue
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(bo[k]), (k, int(b), int(bo[k]))
        elif not fused:
            # BatchNorm running statistics against the oracle after the FIRST step (one forward on identical weights):
            # in-scope layers 5e-3, the PyTorch-run trunk 3e-2 (run-to-run spread of the cuDNN trunk alone is ~1e-2).
            # Later steps are not comparable layer by layer: AdamW's first updates are +-lr per weight whatever the
            # gradient's size, so every weight whose (tiny) gradient changes sign between the fp32 and the bf16 path
            # moves the other way, and this random-init network feeds activations of magnitude ~100 into
            # cross_attention.spatial_reduction - measured (tools/diag_bnstats.py): its running_mean agrees to 2.5e-4
            # after step 1 and differs by 22 % on single channels after step 2 while the losses stay within 0.5 %.
            tol = 3e-2 if k.startswith("pretrained.") else 5e-3
            assert rel_max(bg1[k].float(), bo1[k].float()) <= tol, (k, rel_max(bg1[k].float(), bo1[k].float()))
    assert int(prod.cross_attention.spatial_reduction[1].num_batches_tracked) == 2 * steps


def test_eval_after_replays_sees_current_weights(pkg):
    """train-replay, eval, train-replay, eval - the eager weight-pack cache must never serve stale packs."""
    from depth_b200 import ops
    _, prod = _pair(pkg)
    cfg = fx.loss_config()
    x, t = _batch(2, seed=7)
    opt = _opt(prod)
    with torch.no_grad():
        for g in opt.param_groups:
            g["lr"] = 1e-2                       # large steps: stale packs would be far off
    gstep = pkg.GraphedTrainStep(prod, opt, cfg, x, t, use_rgb=True, world=1, warmup=2)
    for rnd in range(2):
        for _ in range(2):
            gstep(x, t)
        prod.eval()
        with torch.no_grad():
            cached = prod(x)
            ops.PACKS.store.clear()              # force fresh packs: the reference result for this weight state
            fresh = prod(x)
        prod.train()
        assert torch.equal(cached, fresh), f"round {rnd}: eval forward used stale weight packs"
