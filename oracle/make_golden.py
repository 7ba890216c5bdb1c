"""Pin the oracle against the REAL reference and write the golden fixtures under tests/golden/.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden

1. loss / metric functions: reference src/util.py vs oracle/losses.py, bit-for-bit in fp32, on seeded
   full-size inputs (values stored, inputs regenerated from the seed) and on small odd-shaped inputs
   with exact zeros (inputs stored).
2. modules: reference src/network classes vs oracle/model.py with identical deterministic weights;
   outputs, input grads, parameter grads and BN buffers must agree to fp32 round-off; the
   reference's results are stored (sub-sampled) in tests/golden/model_golden.npz.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, fixtures as fx, losses as ol, ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def loss_cases():
    """name -> (pred, target, rgb) generators; the big ones are regenerated from seeds in the tests."""
    out = {}

    def survey():
        torch.manual_seed(0)
        p = torch.rand(4, 1, 448, 576) * 9 + 0.5
        t = torch.rand(4, 1, 448, 576) * 9 + 0.5
        rgb = torch.rand(4, 3, 448, 576)
        return p, t, rgb

    def bench_like():
        g = torch.Generator().manual_seed(1234)
        rgb = torch.randn(3, 3, 448, 576, generator=g)
        t = torch.rand(3, 1, 448, 576, generator=g) * 9.9 + 0.1
        p = t * torch.exp(0.1 * torch.randn(3, 1, 448, 576, generator=g)) * 1.3
        p[:, :, 100:140, 200:300] = 0.0          # exact zeros as the final ReLU produces
        return p, t, rgb

    def small_zeros():
        g = torch.Generator().manual_seed(99)
        rgb = torch.randn(3, 3, 37, 53, generator=g)
        t = torch.rand(3, 1, 37, 53, generator=g) * 5 + 0.2
        p = torch.rand(3, 1, 37, 53, generator=g) * 5
        p[p < 1.0] = 0.0
        t[:, :, 5:9, 7:19] = 0.0                 # invalid target pixels (masked only by SiLog)
        return p, t, rgb

    def tiny():
        g = torch.Generator().manual_seed(5)
        return (torch.rand(1, 1, 2, 3, generator=g) + 0.5, torch.rand(1, 1, 2, 3, generator=g) + 0.5,
                torch.rand(1, 3, 2, 3, generator=g))

    out["survey"] = survey
    out["bench_like"] = bench_like
    out["small_zeros"] = small_zeros
    out["tiny"] = tiny
    return out


THRESH = [1.05, 1.05 ** 2, 1.05 ** 3, 1.25]


def eval_all(mod, p, t, rgb):
    r = {
        "si": mod.scale_invariant_loss(p, t).item(),
        "si_sqrt": mod.scale_invariant_loss(p, t, sqroot=True).item(),
        "silog": mod.silog_loss(p, t, mask=(t > 0)).item(),
        "grad": mod.gradient_loss(p, t).item(),
        "edge": mod.edge_aware_loss(p, t, rgb, 0.5).item(),
        "absrel": mod.absolute_relative_error(p, t).item(),
    }
    for i, th in enumerate(THRESH):
        r[f"delta{i}"] = mod.delta_thres(p, t, thres=th).item()
    return r


def grads_all(mod, p, t, rgb):
    out = {}
    for name, fn in [("si", lambda q: mod.scale_invariant_loss(q, t)),
                     ("silog", lambda q: mod.silog_loss(q, t, mask=(t > 0))),
                     ("grad", lambda q: mod.gradient_loss(q, t)),
                     ("edge", lambda q: mod.edge_aware_loss(q, t, rgb, 0.5))]:
        q = p.clone().requires_grad_(True)
        fn(q).backward()
        out[name] = q.grad
    return out


def do_losses(ref_util):
    gold = {}
    small_store = {}
    for name, gen in loss_cases().items():
        p, t, rgb = gen()
        r_ref = eval_all(ref_util, p, t, rgb)
        r_ora = eval_all(ol, p, t, rgb)
        for k in r_ref:
            a, b = r_ref[k], r_ora[k]
            assert (a == b) or (np.isnan(a) and np.isnan(b)), f"oracle != reference for {name}/{k}: {a} vs {b}"
        r64 = eval_all(ol, p.double(), t.double(), rgb.double())
        counts = ol.delta_counts(p, t, THRESH)
        g_ref = grads_all(ref_util, p, t, rgb)
        g_ora = grads_all(ol, p, t, rgb)
        for k in g_ref:
            a_, b_ = torch.nan_to_num(g_ref[k], 0.0, 0.0, 0.0), torch.nan_to_num(g_ora[k], 0.0, 0.0, 0.0)
            assert float((a_ - b_).abs().max()) <= 1e-6 * float(a_.abs().max()), f"grad mismatch {name}/{k}"
        gold[name] = {"fp32": r_ref, "fp64": r64, "delta_counts": counts.tolist(),
                      "grad_abs_sum": {k: float(torch.nan_to_num(v, 0.0, 0.0, 0.0).abs().double().sum())
                                       for k, v in g_ref.items()}}
        if p.numel() < 10000:
            small_store[f"{name}.pred"] = p.numpy()
            small_store[f"{name}.target"] = t.numpy()
            small_store[f"{name}.rgb"] = rgb.numpy()
            for k, v in g_ref.items():
                small_store[f"{name}.grad_{k}"] = v.numpy()
        print("loss case", name, {k: round(v, 6) for k, v in r_ref.items()})
    # M5 per_pixel_scale_invariant_loss (util.py:159-181): reference vs restatement on one positive image
    gp = torch.Generator().manual_seed(4711)
    t5 = torch.rand(48, 64, generator=gp) * 9.9 + 0.1
    p5 = t5 * torch.exp(0.2 * torch.randn(48, 64, generator=gp)) * 1.3
    m_ref, m_ora = ref_util.per_pixel_scale_invariant_loss(p5, t5), ol.per_pixel_scale_invariant_loss(p5, t5)
    assert torch.equal(m_ref, m_ora), "per_pixel_scale_invariant_loss: oracle != reference"
    gold["per_pixel_si"] = {"sum": float(m_ref.double().sum()), "max": float(m_ref.max())}
    small_store["per_pixel_si.pred"], small_store["per_pixel_si.target"] = p5.numpy(), t5.numpy()
    small_store["per_pixel_si.map"] = m_ref.numpy()
    # main.evaluate_model (main.py:254-392): run the REFERENCE's own function (imported with kornia / omegaconf / wandb
    # stubs) on an identity model over two batches whose prediction resolution differs from the target's, and require
    # the oracle's restatement to reproduce its dict; the inputs and the reference's dict are stored as golden vectors.
    ref_main = ref_import.import_reference_main()
    ev_batches = evaluate_model_case()
    ident = IdentityDepth()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        m_ref = ref_main.evaluate_model(ident, ev_batches, "cpu")
    m_ora = ol.evaluate_model(ident, ev_batches, "cpu")
    for k, v in m_ref.items():
        assert abs(float(v) - float(m_ora[k])) <= 1e-6 * max(abs(float(v)), 1e-12), f"evaluate_model {k}: {v} vs {m_ora[k]}"
    gold["evaluate_model"] = {k: float(v) for k, v in m_ref.items()}
    for i, (x, t, _n) in enumerate(ev_batches):
        small_store[f"evaluate_model.inputs{i}"] = x.numpy()
        small_store[f"evaluate_model.targets{i}"] = t.numpy()
    print("evaluate_model (reference)", {k: round(float(v), 6) for k, v in m_ref.items()})
    p, t, rgb = loss_cases()["small_zeros"]()
    gold["small_zeros"]["evaluate_model_sums"] = ol.evaluate_metric_sums(p, t)
    with open(os.path.join(GOLD, "loss_metric_golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(GOLD, "loss_small_inputs.npz"), **small_store)


class IdentityDepth(nn.Module):
    """'model' whose prediction is channel 0 of its input: drives main.evaluate_model with known depth maps"""

    def forward(self, x):
        return x[:, 0]


def evaluate_model_case():
    """two batches (3 + 2 samples) of (inputs (B,3,30,44) carrying the prediction in channel 0, targets (B,1,24,36), names):
    exact zeros in both prediction and target, values across the 1.25^k thresholds"""
    out = []
    for i, b in enumerate((3, 2)):
        g = torch.Generator().manual_seed(9100 + i)
        t = torch.rand(b, 1, 24, 36, generator=g) * 9.9 + 0.1
        p = torch.nn.functional.interpolate(t, size=(30, 44), mode="bilinear", align_corners=True)
        p = p * torch.exp(0.35 * torch.randn(p.shape, generator=g)) * 1.2
        p[:, :, :2, :5] = 0.0
        t[:, :, 5:7, 3:9] = 0.0
        x = torch.cat([p, torch.rand(b, 2, 30, 44, generator=g)], dim=1)
        out.append((x, t, [f"s{i}_{j}" for j in range(b)]))
    return out


def build_reference(kind, kw, ref_blocks, ref_dpt, ref_sem):
    if kind == "rcu":
        return ref_blocks.ResidualConvUnit_custom(kw["features"], nn.ReLU(False), False)
    if kind == "fusion":
        return ref_blocks.FeatureFusionBlock_custom(kw["features"], nn.ReLU(False), deconv=False, bn=False,
                                                    expand=kw["expand"], align_corners=True)
    if kind == "resblock":
        return ref_sem.ResidualBlock(kw["cin"], kw["cout"])
    if kind == "dinohead":
        return ref_dpt.Dinov2Head(1, 384, 128, use_bn=False, out_channels=[128, 256, 512, 512], use_clstoken=False)
    if kind == "xattn":
        return ref_sem.CrossAttention(kw["dim"], window_size=16)
    if kind == "rcu_large":
        return ref_blocks.ResidualConvUnit(kw["features"])
    if kind == "fusion_large":
        return ref_blocks.FeatureFusionBlock(kw["features"])
    if kind == "dpt":
        # the real DPTDepthModel with an inert backbone object; its own forward() then runs on the four feature maps
        # handed in as `x` (forward_transformer is an instance attribute set in DPT.__init__, dpt_depth.py:214-226)
        real = ref_blocks._make_pretrained_vitb_rn50_384
        ref_blocks._make_pretrained_vitb_rn50_384 = lambda *a, **k: nn.Module()
        try:
            m = ref_dpt.DPTDepthModel(path=None, backbone="vitb_rn50_384", features=kw["features"], non_negative=True)
        finally:
            ref_blocks._make_pretrained_vitb_rn50_384 = real
        m.forward_transformer = lambda pretrained, feats: feats
        return m
    if kind == "midas_large":
        import network.midas_net as ref_large
        real = ref_blocks._make_pretrained_resnext101_wsl
        ref_blocks._make_pretrained_resnext101_wsl = lambda *a, **k: _FeatureFeeder()
        try:
            m = ref_large.MidasNet(None, features=kw["features"], non_negative=True)
        finally:
            ref_blocks._make_pretrained_resnext101_wsl = real
        return m
    raise KeyError(kind)


class _Feed(nn.Module):
    """stands in for one ResNeXt stage: ignores its input and returns the preset feature map"""

    def __init__(self):
        super().__init__()
        self.value = None

    def forward(self, _x):
        return self.value


class _FeatureFeeder(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer1, self.layer2, self.layer3, self.layer4 = _Feed(), _Feed(), _Feed(), _Feed()


def reference_call(kind):
    """how the REAL reference module is driven for the decoder-only cases"""
    if kind == "dpt":
        return lambda m, xs, fkw: m(tuple(xs))
    if kind == "midas_large":
        def call(m, xs, fkw):
            for i, x in enumerate(xs):
                getattr(m.pretrained, f"layer{i + 1}").value = x
            return m(xs[0])
        return call
    return None


def compare(a, b, what, tol=2e-5):
    assert set(a) == set(b), (what, set(a) ^ set(b))
    worst = 0.0
    for k in a:
        scale = max(float(a[k].abs().max()), 1e-6)
        err = float((a[k] - b[k]).abs().max()) / scale
        worst = max(worst, err)
        assert err < tol, f"{what}: {k} rel err {err}"
    return worst


def do_modules(ref_blocks, ref_dpt, ref_sem, ref_small):
    store = {}
    for name, (kind, kw, shapes, fkw) in cases.CASES.items():
        ref = fx.fill_deterministic(build_reference(kind, kw, ref_blocks, ref_dpt, ref_sem))
        ora = fx.fill_deterministic(cases.build_oracle(kind, kw))
        assert list(ref.state_dict().keys()) == list(ora.state_dict().keys()), name
        ora.load_state_dict(ref.state_dict(), strict=True)
        r = cases.run_case(ref, name, call=reference_call(kind))
        o = cases.run_case(ora, name)
        w = compare(r, o, name)
        print(f"module case {name}: oracle vs reference worst rel err {w:.2e} ({len(r)} tensors)")
        for k, v in r.items():
            store[f"{name}/{k}"] = fx.subsample(v).numpy()
    # full models
    standins = fx.load_standins()
    x, t = cases.full_batch()
    for tag, ref_ctor, ora_ctor in [
        ("semantics", lambda: ref_sem.MidasNetSemantics(None, features=64, backbone="efficientnet_lite3",
                                                        exportable=True, non_negative=True, cfg=fx.model_cfg(),
                                                        blocks={"expand": True}, dinov2_type="dinov2_vits14"),
         cases.build_oracle_semantics),
        ("small", lambda: ref_small.MidasNet_small(None, features=64, backbone="efficientnet_lite3", exportable=True,
                                                   non_negative=True, cfg=fx.model_cfg(), blocks={"expand": True}),
         cases.build_oracle_small),
    ]:
        with ref_import.offline_hub(standins.hub_load_standin):
            ref = ref_ctor()
        ora = ora_ctor(standins)
        kr = [k for k in ref.state_dict().keys()]
        ko = [k for k in ora.state_dict().keys()]
        assert kr == ko, (tag, set(kr) ^ set(ko))
        cases.prepare_full(ref)
        ora.load_state_dict(ref.state_dict(), strict=True)
        res = {}
        for m, slot in ((ref, "ref"), (ora, "ora")):
            m.train()
            out = m(x)
            loss = ol.scale_invariant_loss(out.unsqueeze(1), t)
            loss.backward()
            d = {"out": out.detach(), "loss": loss.detach().reshape(1)}
            for k, p in m.named_parameters():
                if p.grad is not None and not k.startswith(("pretrained.", "dinov2.")):
                    d[f"gp.{k}"] = p.grad.detach()
            for k, b in m.named_buffers():
                if not k.startswith(("pretrained.", "dinov2.")):
                    d[f"buf.{k}"] = b.detach().float()
            m.eval()
            with torch.no_grad():
                d["out_eval"] = m(x).detach()
            res[slot] = d
        none_grad = sorted(k for k, p in ref.named_parameters() if p.requires_grad and p.grad is None)
        w = compare(res["ref"], res["ora"], f"full/{tag}", tol=5e-4)
        print(f"full model {tag}: oracle vs reference worst rel err {w:.2e}; loss {float(res['ref']['loss']):.6f}; "
              f"zero fraction {float((res['ref']['out'] == 0).float().mean()):.3f}; params without grad: {len(none_grad)}")
        for k, v in res["ref"].items():
            store[f"full_{tag}/{k}"] = fx.subsample(v, 30000).numpy()
        store[f"full_{tag}/nograd_keys"] = np.array(none_grad)
        store[f"full_{tag}/state_keys"] = np.array(kr)
    np.savez_compressed(os.path.join(GOLD, "model_golden.npz"), **store)


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    ref_util, ref_blocks, ref_dpt, ref_sem, ref_small, ref_large = ref_import.import_reference()
    do_losses(ref_util)
    do_modules(ref_blocks, ref_dpt, ref_sem, ref_small)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
