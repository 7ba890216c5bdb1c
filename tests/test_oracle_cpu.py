"""CPU suite: the oracle against the golden vectors produced from the real reference
(oracle/make_golden.py), and the C-ABI surface of the built library."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import cases, fixtures as fx, losses as ol          # noqa: E402
from oracle.make_golden import loss_cases, THRESH, eval_all    # noqa: E402


@pytest.mark.parametrize("name", ["survey", "bench_like", "small_zeros", "tiny"])
def test_oracle_losses_match_reference_golden(golden_loss, name):
    p, t, rgb = loss_cases()[name]()
    got = eval_all(ol, p, t, rgb)
    exp = golden_loss[name]["fp32"]
    for k, v in exp.items():
        assert got[k] == pytest.approx(v, rel=1e-6, abs=1e-9), (name, k)
    cnt = ol.delta_counts(p, t, THRESH).tolist()
    assert cnt == golden_loss[name]["delta_counts"]


def test_survey_known_answers(golden_loss):
    """the values SURVEY.md section 8c lists, computed by the reference's util.py."""
    g = golden_loss["survey"]["fp32"]
    assert g["si"] == pytest.approx(0.982784748, rel=2e-6)
    assert g["si_sqrt"] == pytest.approx(0.991354585, rel=2e-6)
    assert g["silog"] == pytest.approx(0.982787371, rel=2e-6)
    assert g["grad"] == pytest.approx(4.79985094, rel=2e-6)
    assert g["edge"] == pytest.approx(1.06726217, rel=2e-6)
    assert g["absrel"] == pytest.approx(1.08991563, rel=2e-6)
    assert g["delta0"] == pytest.approx(0.0532449372, rel=2e-6)
    assert g["delta3"] == pytest.approx(0.222841293, rel=2e-6)


def test_small_stored_inputs_roundtrip(golden_loss):
    z = np.load(os.path.join(ROOT, "tests", "golden", "loss_small_inputs.npz"))
    p, t, rgb = (torch.from_numpy(z[f"small_zeros.{k}"]) for k in ("pred", "target", "rgb"))
    got = eval_all(ol, p, t, rgb)
    for k, v in golden_loss["small_zeros"]["fp32"].items():
        assert got[k] == pytest.approx(v, rel=1e-6, abs=1e-9)


@pytest.mark.parametrize("name", ["rcu64", "fusion128_expand", "resblock_64_32", "xattn_multi", "dinohead", "rcu_large64",
                                  "fusion_large64", "fusion_large64_single", "dpt_decoder64", "midas_large_decoder64"])
def test_oracle_modules_match_reference_golden(name):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))
    kind, kw, shapes, fkw = cases.CASES[name]
    m = fx.fill_deterministic(cases.build_oracle(kind, kw))
    res = cases.run_case(m, name)
    for k, v in res.items():
        ref = torch.from_numpy(gold[f"{name}/{k}"])
        got = fx.subsample(v)
        scale = max(float(ref.abs().max()), 1e-6)
        assert float((got - ref).abs().max()) / scale < 1e-4, (name, k)


def test_cross_attention_last_writer_identity():
    """The closed form the CUDA kernel uses: token i attends over the key range of the LAST window containing i."""
    from oracle.model import CrossAttention
    hr, wr, ws = 20, 24, 16
    rng = CrossAttention.window_ranges(hr, wr, ws)
    owner = np.full(hr * wr, -1)
    for wi, (lo, hi) in enumerate(rng):
        owner[lo:hi] = wi
    assert (owner >= 0).all()
    # default geometry (56x72): SURVEY section 8 A9 owner histogram
    rng = CrossAttention.window_ranges(56, 72, 16)
    owner = np.full(56 * 72, -1)
    for wi, (lo, hi) in enumerate(rng):
        owner[lo:hi] = wi
    hist = np.bincount(owner, minlength=20)
    assert hist.tolist() == [16, 16, 16, 16, 1088] * 3 + [16, 16, 16, 16, 512]
    total_scores = sum(hist[w] * (rng[w][1] - rng[w][0]) for w in range(20))
    assert total_scores < 4.4e6


def test_library_exports_every_declared_symbol():
    """every function include/depth_b200.h declares is exported by the built .so (no compute calls)."""
    import depth_b200
    hdr = open(os.path.join(ROOT, "include", "depth_b200.h")).read()
    names = sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 10
    lib = ctypes.CDLL(depth_b200._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.dp_abi_version() == 3


def test_no_cpu_fallback():
    import depth_b200
    p = torch.rand(1, 1, 4, 4) + 0.5
    with pytest.raises(depth_b200._lib.DepthB200Error):
        depth_b200.scale_invariant_loss(p, p)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "monocular-depth-estimation-cil_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, re.M), f"{f} imports the oracle"
                assert "oracle/" not in src and "oracle." not in src, f"{f} references the oracle"


def test_load_model_reads_reference_checkpoint_formats(tmp_path):
    """util.load_model (reference util.py:222-238): both checkpoint layouts main.py / DataParallel runs produce load
    into the drop-in model with identical keys (no CUDA needed: only parameter containers are touched)."""
    import types
    import torch
    import depth_b200
    from depth_b200 import standins
    from depth_b200.network import blocks
    blocks.hub_load = standins.hub_load_standin
    cfg = types.SimpleNamespace(use_lb=False, use_dgr=False, dinov2_type="dinov2_vits14")
    from depth_b200.network.midas_semantics import MidasNetSemantics
    torch.manual_seed(0)
    src = MidasNetSemantics(None, features=64, backbone="efficientnet_lite3", exportable=True, non_negative=True, cfg=cfg,
                            blocks={'expand': True}, dinov2_type="dinov2_vits14")
    p1, p2 = tmp_path / "a.pth", tmp_path / "b.pth"
    torch.save({"epoch": 3, "model_state_dict": src.state_dict()}, p1)
    torch.save({"module." + k: v for k, v in src.state_dict().items()}, p2)
    for p in (p1, p2):
        m = depth_b200.util.load_model("MiDaS_small", str(p), cfg)
        sd = m.state_dict()
        assert list(sd.keys()) == list(src.state_dict().keys())
        assert all(torch.equal(sd[k], v) for k, v in src.state_dict().items())


def test_attention_segments_match_window_loop():
    """Host logic of the last-writer attention (ops.attention_segments): every token's key range equals the range of the
    LAST window of the reference loop (midas_semantics.py:93-112) that contains it, and the 32-query work items tile the
    query runs exactly."""
    import depth_b200
    from depth_b200 import ops
    for hr, wr, ws in [(56, 72, 16), (8, 12, 16), (20, 24, 16), (7, 9, 4)]:
        items, segs, unowned = ops.attention_segments(hr, wr, ws, "cpu")
        N = hr * wr
        owner = [None] * N
        for h in range((hr + ws - 1) // ws):
            for w in range((wr + ws - 1) // ws):
                lo = h * ws * wr + w * ws
                hi = min(min(h * ws + ws, hr) * wr + min(w * ws + ws, wr), N)
                for i in range(lo, hi):
                    owner[i] = (lo, hi)
        got = [None] * N
        for q0, nq, klo, khi in items.tolist():
            assert 1 <= nq <= 32
            for i in range(q0, q0 + nq):
                assert got[i] is None, "work items must not overlap"
                got[i] = (klo, khi)
        assert got == owner
        assert unowned == [i for i in range(N) if owner[i] is None]
        covered = sorted(i for a, b, _, _ in segs.tolist() for i in range(a, b))
        assert covered == [i for i in range(N) if owner[i] is not None]


def test_fused_encoder_structure_recognition():
    """network/encoder_fused.supported(): the gen-efficientnet-shaped trunk is accepted; anything it cannot execute
    exactly (biased conv, squeeze-excite, unknown block) sends the caller back to the PyTorch path."""
    import torch.nn as nn
    import depth_b200
    from depth_b200 import standins
    from depth_b200.network import blocks, encoder_fused
    blocks.hub_load = standins.hub_load_standin
    p = blocks._make_pretrained_efficientnet_lite3(False)
    assert encoder_fused.supported(p)
    geo = encoder_fused._dw_geom(p.layer2[0][0].conv_dw, 112, 144)
    assert geo == (2, 2, 2, 56, 72)
    q = blocks._make_pretrained_efficientnet_lite3(False)
    q.layer3[0][1].conv_pw = nn.Conv2d(96, 576, 1, bias=True)
    assert not encoder_fused.supported(q)
    r = blocks._make_pretrained_efficientnet_lite3(False)
    r.layer2[0][0].se = nn.Sequential(nn.Conv2d(8, 8, 1))
    assert not encoder_fused.supported(r)
    s = blocks._make_pretrained_efficientnet_lite3(False)
    s.layer4[0][0] = nn.Sequential(nn.Conv2d(136, 232, 3, 2, 1), nn.ReLU())
    assert not encoder_fused.supported(s)


def test_eval_plan_covers_every_sample_shape():
    """dp_eval_metrics_plan (host arithmetic of the streaming evaluation kernel's decomposition): for a sweep of image
    sizes on a B200-like device (228 KB of shared memory per SM, 148 SMs) the slices tile the sample, are 16-byte
    aligned, fit two slots of shared memory per CTA, and the groups fit the SM count; shapes whose pixel count is not a
    multiple of 4 are declined (they run the cluster kernel)."""
    import depth_b200
    lib = ctypes.CDLL(depth_b200._lib.LIB_PATH)
    lib.dp_eval_metrics_plan.restype = ctypes.c_int
    lib.dp_eval_metrics_plan.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_size_t)]
    smem, sms = 233472, 148
    shapes = [(448, 576), (426, 560), (896, 1152), (64, 96), (32, 40), (2, 2), (1080, 1920), (480, 640), (37, 52)]
    for H, W in shapes:
        for B, nthr in ((1, 3), (3, 1), (32, 2), (650, 3), (650, 8)):
            G, groups, per, dyn = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
            ok = lib.dp_eval_metrics_plan(H * W, B, nthr, smem, sms, ctypes.byref(G), ctypes.byref(groups),
                                          ctypes.byref(per), ctypes.byref(dyn))
            assert ok == 1, (H, W, B)
            static = 2560 if nthr in (1, 3) else 3584      # ptxas: 2432 / 3456 bytes of static shared memory
            n = H * W
            assert per.value % 4 == 0 and per.value * G.value >= n           # slices tile the sample, 16-byte aligned
            assert per.value * (G.value - 1) < n + 4 * G.value                 # no CTA beyond the ragged last one is idle
            assert 1 <= groups.value <= B and groups.value * G.value <= sms    # every CTA of every group is resident
            assert dyn.value == per.value * 8 * 2 and dyn.value + static + 1024 <= smem   # static + reserved smem fit
    assert lib.dp_eval_metrics_plan(37 * 53, 4, 3, smem, sms, None, None, None, None) == 0      # odd pixel count
    assert lib.dp_eval_metrics_plan(448 * 576, 4, 3, 16384, sms, None, None, None, None) == 0   # no room for a slot
    assert lib.dp_eval_metrics_plan(8000 * 8000, 4, 3, smem, sms, None, None, None, None) == 0  # sample larger than the chip's smem


def test_evaluate_model_oracle_reproduces_reference_golden(golden_loss):
    """M4: oracle.losses.evaluate_model vs the dict the REFERENCE's main.evaluate_model produced on the stored inputs
    (oracle/make_golden.py runs the reference's own function: identity model, two batches, 30x44 predictions against
    24x36 targets)."""
    import os
    import numpy as np
    import torch
    from oracle import losses as ol
    store = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_small_inputs.npz"))

    class Ident(torch.nn.Module):
        def forward(self, x):
            return x[:, 0]

    batches = [(torch.from_numpy(store[f"evaluate_model.inputs{i}"]), torch.from_numpy(store[f"evaluate_model.targets{i}"]), None)
               for i in range(2)]
    got = ol.evaluate_model(Ident(), batches, "cpu")
    for k, v in golden_loss["evaluate_model"].items():
        assert abs(got[k] - v) <= 1e-6 * max(abs(v), 1e-12), (k, got[k], v)


def test_frozen_cast_is_transparent_on_the_host():
    """network/frozen_cast.py on the CPU side: only frozen nn.Linear modules get the caching forward, the state_dict and
    deepcopy keep working, and without CUDA autocast the forward is the plain fp32 F.linear (no copies are made)."""
    import copy
    import torch
    import torch.nn as nn
    import depth_b200  # noqa: F401
    from depth_b200.network import frozen_cast
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(8, 16), nn.LayerNorm(16), nn.Linear(16, 4, bias=False), nn.Linear(4, 4))
    for p in list(m[0].parameters()) + list(m[2].parameters()):
        p.requires_grad_(False)
    ref = copy.deepcopy(m)
    keys = list(m.state_dict().keys())
    frozen_cast.enable(m)
    assert "forward" in m[0].__dict__ and "forward" in m[2].__dict__ and "forward" not in m[3].__dict__
    x = torch.randn(5, 8)
    assert torch.equal(m(x), ref(x))
    assert "_dp_cast" not in m[0].__dict__                      # no autocast, no CUDA: nothing cached
    assert list(m.state_dict().keys()) == keys
    m2 = copy.deepcopy(m)
    assert m2[0].forward.__self__ is m2[0] and torch.equal(m2(x), ref(x))
    frozen_cast.refresh_all()                                   # nothing to refresh; must not raise
