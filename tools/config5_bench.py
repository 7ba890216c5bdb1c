"""BASELINE config 5 on one GPU: the largest configured decoder (DPT, features=256; reference dpt_depth.py:155-293) at
2x input resolution (896x1152), bf16, fed synthetic encoder maps [256, 512, 768, 768] at strides 4..32 (the timm
backbone is third-party: SURVEY 8c).  One step = decoder forward + SI loss + backward (parameters and feature maps),
timed with CUDA events.  Algorithmic work: 807.2 GFLOP/img forward, x3 for the train step.
usage: python tools/config5_bench.py [batch=8] [steps=5] [layers]   (layers: per-shape tcgen05 conv / wgrad timings)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_b200
from depth_b200.network import dpt_depth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
H, W = 896, 1152
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = dpt_depth.DPTDepthModel(path=None, backbone="vitb_rn50_384", features=256, non_negative=True).to(dev).train()
with torch.no_grad():      # keep the head's ReLU alive on random weights
    for k, p in model.named_parameters():
        if k.endswith("output_conv.4.bias") or k.endswith("head.4.bias"):
            p.fill_(1.0)
g = torch.Generator(device=dev).manual_seed(1234)
feats = [torch.randn(B, c, H // s, W // s, device=dev, generator=g).requires_grad_(True)
         for c, s in zip((256, 512, 768, 768), (4, 8, 16, 32))]
target = torch.rand(B, 1, H, W, device=dev, generator=g) * 9.9 + 0.1


def step():
    for p in model.parameters():
        p.grad = None
    for f in feats:
        f.grad = None
    out = model.forward_features(*feats)
    loss = depth_b200.scale_invariant_loss(out.unsqueeze(1), target)
    loss.backward()
    return loss


for _ in range(3):
    loss = step()
torch.cuda.synchronize()
n0 = depth_b200._lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ips = B / (ms / 1e3)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
tf = 3 * 807.2e9 * ips / 1e12
peak = float(peaks.get("bf16_tflops_sustained", 1383.7)) if isinstance(peaks, dict) else 1383.7
print(json.dumps({"workload": "configs[4] on 1 GPU: DPT decoder features=256 @896x1152, bf16, fwd + SI loss + bwd (eager launches)",
                  "batch": B, "steps": steps, "ms_per_step": round(ms, 2), "images_per_s": round(ips, 2),
                  "algorithmic_tflops": round(tf, 1), "frac_of_sustained_bf16_peak": round(tf / peak, 3),
                  "gpu_launches_per_step": (depth_b200._lib.launch_count() - n0) // steps,
                  "loss": float(loss.item()), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}))

if len(sys.argv) > 3 and sys.argv[3] == "layers":
    from depth_b200 import ops
    rec = []
    orig_conv, orig_wg = ops._conv_tc_launch, ops._wgrad_tc

    def conv_hook(x, wp, Cout, KS, *rest):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = orig_conv(x, wp, Cout, KS, *rest); e.record()
        Bq, Hq, Wq, Cin = x.shape
        rec.append(("conv", (Hq, Wq, Cin, Cout, KS), 2.0 * Bq * Hq * Wq * Cin * Cout * KS * KS, s, e))
        return r

    def wg_hook(x, g_, Cin, Cout, KS):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = orig_wg(x, g_, Cin, Cout, KS); e.record()
        Bq, Hq, Wq, _ = x.shape
        rec.append(("wgrad", (Hq, Wq, Cin, Cout, KS), 2.0 * Bq * Hq * Wq * Cin * Cout * KS * KS, s, e))
        return r

    ops._conv_tc_launch, ops._wgrad_tc = conv_hook, wg_hook
    torch.cuda._sleep(int(0.2 * 1.9e9))      # let the host run ahead: events bracket GPU time only
    step()
    torch.cuda.synchronize()
    ops._conv_tc_launch, ops._wgrad_tc = orig_conv, orig_wg
    per = {}
    for kind, shp, f, s, e in rec:
        d = per.setdefault((kind,) + shp, [0.0, 0.0, 0])
        d[0] += f; d[1] += s.elapsed_time(e); d[2] += 1
    tot = sum(d[1] for d in per.values())
    print(f"tcgen05 launches: {tot:.2f} ms of {ms:.2f} ms/step, {sum(d[0] for d in per.values()) / tot / 1e9:.0f} TFLOP/s aggregate")
    for k, d in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{str(k):46s} n={d[2]:3d} {d[1]:8.3f} ms {d[0] / d[1] / 1e9:8.1f} TFLOP/s")
    from torch.profiler import profile, ProfilerActivity
    import collections
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0:
            agg[ev.name][0] += ev.device_time; agg[ev.name][1] += 1
    tk = sum(v[0] for v in agg.values())
    print(f"kernel time {tk / 1e3:.2f} ms")
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"{us / 1e3:9.3f} ms {100 * us / tk:5.1f}%  n={n:4d}  {name[:110]}")
