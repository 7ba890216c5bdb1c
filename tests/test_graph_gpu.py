"""GraphedTrainStep (whole train step in one CUDA graph): replay reproduces the eager step, the side-stream weight
gradients are bit-identical to the single-stream step, and the fused EfficientNet trunk trains through the graph."""
import copy

import pytest
import torch

from oracle import cases, fixtures as fx

pytestmark = pytest.mark.gpu


def _model(pkg, fused_encoder):
    from depth_b200 import standins
    from depth_b200.network import blocks, midas_semantics
    blocks.hub_load = standins.hub_load_standin
    torch.manual_seed(0)
    m = midas_semantics.MidasNetSemantics(None, features=64, backbone="efficientnet_lite3", exportable=True,
                                          non_negative=True, cfg=fx.model_cfg(), blocks={'expand': True},
                                          dinov2_type='dinov2_vits14')
    with torch.no_grad():
        m.depth_head[1].bias.add_(2.0)
    m.fused_encoder = fused_encoder
    return m.cuda().train()


def _run(pkg, model, side, steps=2):
    x, t = cases.full_batch()
    x, t = x.cuda(), t.cuda()
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, fused=True,
                            capturable=True)
    g = pkg.GraphedTrainStep(model, opt, fx.loss_config(), x, t, use_rgb=True, world=1, warmup=2, side_wgrad=side)
    for _ in range(steps):
        g()
    torch.cuda.synchronize()
    g.finish()
    return g.loss_dict(), {k: v.detach().clone() for k, v in model.state_dict().items()}


@pytest.mark.parametrize("fused_encoder", [False, True])
def test_side_stream_weight_gradients_bit_identical(pkg, fused_encoder):
    base = _model(pkg, fused_encoder)
    a, b = copy.deepcopy(base), copy.deepcopy(base)
    la, sa = _run(pkg, a, side=False)
    lb, sb = _run(pkg, b, side=True)
    assert la == lb, (la, lb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    # the step actually trained: parameters moved and the loss is finite
    moved = sum(int(not torch.equal(sa[k], v)) for k, v in base.state_dict().items() if v.dtype.is_floating_point)
    assert moved > 100 and all(map(lambda v: v == v, la.values()))
