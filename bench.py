#!/usr/bin/env python
"""bench.py - the driver's measurement contract.

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

Metric (BASELINE.json): train images/s of the default model (MidasNetSemantics, features 64, efficientnet_lite3 +
dinov2_vits14 stand-ins, 448x576), forward + combined loss (config.yaml weights 1/0/0/0) + backward + AdamW, bf16
activations, batch 32 per GPU (weak scaling).  `value` times steps whose inputs are already in HBM; `e2e` times the
same steps fed from pinned host memory through the public API with a device->host read of the loss.  Extra keys:
roofline (tcgen05 conv kernels, measured live with CUDA events), cpu_baseline (oracle port on the host cores),
eval (fused SI-RMSE/AbsRel/delta reductions in Gpx/s against the HBM roofline).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 448, 576            # INPUT_SIZE, reference src/main.py:31
METRIC = "train images/s (fwd+bwd+SI loss+AdamW), default MidasNetSemantics 448x576"
FWD_GFLOP_PER_IMG = 113.63 + 2.80      # in-repo conv/linear + attention matmuls, reference formulation


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-overlap-h2d", action="store_true", help="copy each batch on the main stream instead of under the previous replay")
    ap.add_argument("--profile-layers", action="store_true", help="print per-shape conv timings to stderr")
    ap.add_argument("--no-side-wgrad", action="store_true", help="keep weight gradients on the main stream")
    ap.add_argument("--torch-encoder", action="store_true", help="run the EfficientNet-Lite3 trunk through PyTorch/cuDNN")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_model(device, fused_encoder=True):
    import torch
    import depth_b200  # noqa: F401
    from depth_b200 import standins
    from depth_b200.network import blocks, midas_semantics
    from depth_b200 import config as fx
    blocks.hub_load = standins.hub_load_standin
    torch.manual_seed(0)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = midas_semantics.MidasNetSemantics(None, features=64, backbone="efficientnet_lite3", exportable=True,
                                                  non_negative=True, cfg=fx.model_cfg(), blocks={'expand': True},
                                                  dinov2_type='dinov2_vits14')
    with torch.no_grad():
        model.depth_head[1].bias.add_(2.0)       # keep the ReLU'd random-init depth away from all-zero
    model.encoder_autocast = True      # the frozen DINOv2 stand-in (third-party) runs under bf16 autocast in PyTorch
    model.fused_encoder = fused_encoder  # EfficientNet-Lite3 trunk on libdepth_b200.so (network/encoder_fused.py)
    return model.to(device).train()


def synthetic_batch(B, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, H, W, generator=g)
    t = torch.rand(B, 1, H, W, generator=g) * 9.9 + 0.1
    return x, t


def _timeit(fn, reps, barrier, dev, world, dist):
    """CUDA-event time of `reps` calls of fn() in ms per call, max over ranks"""
    import torch
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        q = torch.tensor([ms], device=dev)
        dist.all_reduce(q, op=dist.ReduceOp.MAX)
        ms = float(q.item())
    return ms


def _profile_traffic(name, files, index=0, kernel=None):
    """`traffic` (DRAM bytes per launch) from a committed ncu --set full summary (tools/ncu_summary.py), paired with a
    live timing only if the capture was taken from the same kernel sources: the summary carries the sha1 of every csrc
    file, and the files the kernel is compiled from must match the library that is running."""
    import depth_b200
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            cap = json.load(f)
    except Exception:
        return None, f"profiles/{name} not found"
    now = depth_b200._lib.build_files()
    stale = [f for f in files if cap.get("source_files", {}).get(f) != now.get(f)]
    if stale or not now:
        return None, f"profiles/{name} was captured from other sources ({', '.join(stale) or 'no build record'} changed): not paired"
    ls = [k for k in cap["launches"] if kernel is None or kernel in k["kernel"]]
    k = ls[index]
    return int(k["dram_bytes_read"] + k["dram_bytes_write"]), \
        f"profiles/{name} launch '{k['kernel'][:40]}' (dram__bytes_read.sum + dram__bytes_write.sum, one launch; sources unchanged since the capture)"


def eval_kernel_leg(depth_b200, dev, rank, world, barrier, dist, hbm, build):
    """second half of the BASELINE metric: the fused evaluation reductions, Gpx/s against the HBM roofline"""
    import torch
    EB = 650          # test-set-like sample count, sharded by sample across ranks
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    tt = torch.rand(EB, 1, H, W, device=dev, generator=g) * 9.9 + 0.1
    pp = tt * torch.exp(0.1 * torch.randn(EB, 1, H, W, device=dev, generator=g)) * 1.3
    pp[:, :, 100:140, 200:300] = 0.0
    res = {}
    for key, fm in (("default", None), ("exact", False), ("mufu", True)):
        for _ in range(3):
            depth_b200.evaluation_metrics(pp, tt, fast_math=fm)
        ms = _timeit(lambda: depth_b200.evaluation_metrics(pp, tt, fast_math=fm), 10, barrier, dev, world, dist)
        res[key] = (ms, depth_b200.evaluation_metrics(pp, tt, fast_math=fm).tolist())
    px = EB * H * W
    ms = res["default"][0]
    # the contract, checked on the benched inputs: default vs the exact IEEE path
    d, e = res["default"][1], res["exact"][1]
    contract = {"si_rmse_rel_diff": abs(d[0] - e[0]) / abs(e[0]), "abs_rel_rel_diff": abs(d[1] - e[1]) / abs(e[1]),
                "max_delta_fraction_diff": max(abs(x - y) for x, y in zip(d[2:], e[2:])),
                "allowed": "1e-5 relative / 1e-4 (0.01 % of pixels)"}
    traffic, tsrc = _profile_traffic("ncu_kernels_r2.json", ["loss_metrics.cu", "tc.cuh", "common.cuh"], kernel="eval_stream_kernel")

    def roof(ms_):
        gbs = px * 8 / (ms_ / 1e3) / 1e9
        return {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm, "unit": "GB/s", "frac": round(gbs / hbm, 4)}

    r = roof(ms)
    r.update({"algorithmic_bytes_per_px": 8, "algorithmic_bytes_per_launch": px * 8, "traffic": traffic, "traffic_source": tsrc})
    del pp, tt
    return {"metric": "eval Gpx/s (fused SI-RMSE + AbsRel + 3x delta, evaluation.py:157-166)",
            "value": round(world * px / (ms / 1e3) / 1e9, 2), "unit": "Gpx/s", "batch_per_gpu": EB,
            "inputs": f"{px * 8 / 1e6:.0f} MB per GPU per call (> 126 MB L2)",
            "kernel": "eval_stream_kernel: one CTA per SM, groups of CTAs own a sample, slices staged in shared memory by "
                      "cp.async.bulk (4 slots of one chunk each), per-sample scale exchanged through global memory "
                      "by a pipelined warp, classification two samples behind the moments sweep, from shared memory",
            "arithmetic": "default path of evaluation_metrics: one shared reciprocal + one lg2 per pixel, division-free "
                          "threshold test (exact code for slices with negative values / non-finite scales)",
            "contract_vs_exact_on_these_inputs": contract, "roofline": r,
            "exact": {"value": round(world * px / (res["exact"][0] / 1e3) / 1e9, 2), "unit": "Gpx/s",
                      "arithmetic": "IEEE logf / division, the reference's own arithmetic (the checker)", "roofline": roof(res["exact"][0])},
            "mufu": {"value": round(world * px / (res["mufu"][0] / 1e3) / 1e9, 2), "unit": "Gpx/s",
                     "arithmetic": "round-1 MUFU lg2 / rcp per operand", "roofline": roof(res["mufu"][0])}}


def eval_loop_leg(depth_b200, model, dev, rank, world, barrier, dist):
    """BASELINE configs[3]: evaluation.py's metric loop + generate_predictions' writer over a synthetic test set,
    sharded by sample over the ranks: eval-mode forward of the default model, fused metric kernel per batch, one
    all-reduce of the partial sums (NCCL), then the prediction pass (forward, 426x560 resize, .npy files to tmpfs)."""
    import shutil
    import tempfile
    import torch
    N_PER_RANK, BS = 160, 32
    g = torch.Generator().manual_seed(99 + rank)
    xs = torch.randn(N_PER_RANK, 3, H, W, generator=g).pin_memory()
    ts = (torch.rand(N_PER_RANK, 1, H, W, generator=g) * 9.9 + 0.1).pin_memory()
    names = [f"x x_{rank}_{i:05d}.npy" for i in range(N_PER_RANK)]

    shard = [(xs[lo:lo + BS], ts[lo:lo + BS], names[lo:lo + BS]) for lo in range(0, N_PER_RANK, BS)]

    def metric_pass():
        return depth_b200.evaluation.evaluate_batches(model, shard, dev, local_shard=True)

    out = metric_pass()                       # warm-up (allocator pools, eval-mode weight packs)
    REPS = 3                                  # a pass is ~0.1-0.3 s of host-driven work: median of three
    tms = []
    for _ in range(REPS):
        barrier()
        t0 = time.perf_counter()
        out = metric_pass()
        barrier()
        tms.append(time.perf_counter() - t0)
    t_metric = sorted(tms)[REPS // 2]
    loader = [(xs[lo:lo + BS], names[lo:lo + BS]) for lo in range(0, N_PER_RANK, BS)]
    tps = []
    nfiles = 0
    for _ in range(REPS):
        tmp = tempfile.mkdtemp(prefix="dp_pred_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            barrier()
            t0 = time.perf_counter()
            depth_b200.util.generate_test_predictions(model, loader, dev, tmp)
            barrier()
            tps.append(time.perf_counter() - t0)
            nfiles = len(os.listdir(tmp))
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    t_pred = sorted(tps)[REPS // 2]
    model.train()
    tot = world * N_PER_RANK
    return {"workload": f"configs[3]: evaluation.py metric loop + generate_predictions writer, {N_PER_RANK} synthetic samples "
                        f"per GPU (batch {BS}, host-resident pinned inputs, H2D inside the timed region), sharded over {world} rank(s)",
            "metric_pass": {"images_per_s": round(tot / t_metric, 1), "seconds": round(t_metric, 3),
                            "what": "eval-mode forward + eval_stream_kernel per batch + one all-reduce of 6 doubles"},
            "prediction_pass": {"images_per_s": round(tot / t_pred, 1), "seconds": round(t_pred, 3), "files_per_rank": nfiles,
                                "what": "eval-mode forward + fp32 resize to 426x560 + one D2H per batch + np.save to tmpfs"},
            "samples": out["samples"],
            "metrics": {"si_rmse": out["si_rmse"], "abs_rel": out["abs_rel"], "delta": out["delta"]},
            "timing": "host wall clock around the whole pass (includes the per-batch H2D copies and the final D2H), max over "
                      "ranks by barrier; median of 3 passes after one warm-up pass",
            "seconds_all": {"metric_pass": [round(v, 3) for v in tms], "prediction_pass": [round(v, 3) for v in tps]}}


def config5_leg(depth_b200, dev, rank, world, barrier, dist, tf_sus):
    """BASELINE configs[4]: the largest configured decoder (DPT, features=256, dpt_depth.py:155-293) at 2x input
    resolution (896x1152), bf16, 8 images per GPU, fed synthetic encoder maps [256,512,768,768] at strides 4..32 (the timm
    backbone is third-party).  One step = decoder forward + SI loss + backward + gradient all-reduce + AdamW."""
    import torch
    from depth_b200.network import dpt_depth
    from depth_b200 import distributed as D
    B5, H5, W5 = 8, 896, 1152
    torch.manual_seed(0)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = dpt_depth.DPTDepthModel(path=None, backbone="vitb_rn50_384", features=256, non_negative=True).to(dev).train()
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.endswith("output_conv.4.bias") or k.endswith("head.4.bias"):
                p.fill_(1.0)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    feats = [torch.randn(B5, c, H5 // s, W5 // s, device=dev, generator=g).requires_grad_(True)
             for c, s in zip((256, 512, 768, 768), (4, 8, 16, 32))]
    target = torch.rand(B5, 1, H5, W5, device=dev, generator=g) * 9.9 + 0.1
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4, fused=True)
    red = D.GradientAllReducer(params, world)

    def step():
        red.zero()
        for f in feats:
            f.grad = None
        out = model.forward_features(*feats)
        loss = depth_b200.scale_invariant_loss(out.unsqueeze(1), target)
        loss.backward()
        red.reduce()
        opt.step()

    for _ in range(3):
        step()
    n0 = depth_b200._lib.launch_count()
    ms = _timeit(step, 4, barrier, dev, world, dist)
    launches = (depth_b200._lib.launch_count() - n0) // 4
    ips = world * B5 / (ms / 1e3)
    tf = 3 * 807.2e9 * (ips / world) / 1e12
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    del model, feats, target, opt, red
    torch.cuda.empty_cache()
    return {"workload": f"configs[4]: DPT decoder features=256 @896x1152, bf16, batch {B5}/GPU, fwd + SI loss + bwd + "
                        "gradient all-reduce + AdamW (eager launches), synthetic encoder maps",
            "value": round(ips, 2), "unit": "images/s", "ms_per_step": round(ms, 2), "gpu_launches_per_step": int(launches),
            "roofline": {"bound": "tensor", "achieved": round(tf, 1), "peak": tf_sus, "unit": "TFLOP/s per GPU",
                         "frac": round(tf / tf_sus, 4),
                         "algorithmic_flops_per_image": 3 * 807.2e9, "what": "whole step against the sustained bf16 peak"},
            "peak_mem_gb": round(mem, 1)}


def torch_gpu_baseline(B, steps):
    """BASELINE leg, not the product: the reference's modules (oracle port, pinned to the reference) run by stock PyTorch
    on the same B200 - bf16 autocast, channels_last, cuDNN / cuBLAS, fused AdamW - on the same step and batch.  This is
    the bar SURVEY fact 1 names (PyTorch-eager on B200); the CPU baseline says nothing about kernel quality."""
    import torch
    from oracle import cases, fixtures as fx, losses as ol
    standins = fx.load_standins()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = cases.build_oracle_semantics(standins)
    with torch.no_grad():
        model.depth_head[1].bias.add_(2.0)
    model = model.to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, fused=True)
    x, t = synthetic_batch(B, 1234)
    x = x.to(dev).contiguous(memory_format=torch.channels_last)
    t = t.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x)
        loss = ol.scale_invariant_loss(out.float().unsqueeze(1), t)
        loss.backward()
        opt.step()
        return loss

    try:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res = {"value": round(B / (ms / 1e3), 2), "unit": "images/s", "ms_per_step": round(ms, 2), "batch": B, "steps": steps,
               "kind": "port", "what": "oracle modules under torch.autocast(bf16) + channels_last, cuDNN/cuBLAS, eager, "
                                       f"fused AdamW; torch {torch.__version__}", "loss": float(loss.item())}
    except torch.OutOfMemoryError as ex:      # report, never fail the bench on the baseline
        res = {"unavailable": f"out of memory at batch {B}: {str(ex)[:80]}"}
    del model, opt
    torch.cuda.empty_cache()
    return res


def run_ours(a):
    import torch
    import torch.distributed as dist
    import depth_b200
    from depth_b200 import distributed as D, ops
    from depth_b200 import config as fx                   # the product arm never imports oracle/
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")            # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    rank, local, world = D.init_from_env()
    assert world == a.gpus or world == 1 and a.gpus == 1, f"launch with torchrun --nproc-per-node {a.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = a.batch
    build = depth_b200._lib.build_id()
    if os.environ.get("DP_NO_BN_FUSION"):
        ops.Fusion.prologue, ops.Fusion.backward = False, False
    model = build_model(dev, fused_encoder=not a.torch_encoder)
    cfg = fx.loss_config()                                # config.yaml:34-42 -> 1 / 0 / 0 / 0
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, fused=True,
                            capturable=True)
    xh, th = synthetic_batch(B, 1234 + rank)
    xh, th = xh.pin_memory(), th.pin_memory()
    xd, td = xh.to(dev), th.to(dev)
    # The whole step (forward, combined_loss, backward, NCCL gradient all-reduce, AdamW) is one CUDA graph.
    gstep = depth_b200.GraphedTrainStep(model, opt, cfg, xd, td, use_rgb=True, world=world, warmup=max(a.warmup, 3),
                                        side_wgrad=not a.no_side_wgrad, overlap_h2d=not a.no_overlap_h2d)
    red = gstep.red

    def step(x, t, read_loss):
        """the same step issued eagerly through the reference-shaped API (model(x), combined_loss, backward, step)"""
        red.zero()
        out = model(x).unsqueeze(1)
        loss, parts = depth_b200.combined_loss(out, t, cfg, rgb=x)      # one fused pass + one D2H read of 8 floats
        loss.backward()
        red.reduce()
        opt.step()
        return parts["si_loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, from_host, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(nsteps):
            last = fn(from_host)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, last

    e2e_first = [0]

    def graph_step(from_host):
        if from_host:
            first = gstep._step == e2e_first[0]
            gstep(xh, th)                       # H2D of this step's batch from pinned memory
            # every step's loss scalars are read back from pinned memory; the read trails the device by one step so
            # that the next batch's H2D overlaps the current replay (the last step's loss is read after the loop)
            return gstep.loss_dict(lag=0 if first else 1)["si_loss"]
        gstep()
        return None

    def eager_step(from_host):
        if from_host:
            return step(xh.to(dev, non_blocking=True), th.to(dev, non_blocking=True), True)
        return step(xd, td, True)

    for _ in range(max(a.warmup, 3)):
        gstep()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, _ = timed(a.steps, False, graph_step)
    clocks = sampler.stop() if rank == 0 else None
    e2e_first[0] = gstep._step
    ms_e2e, _ = timed(a.steps, True, graph_step)
    last_loss = gstep.loss_dict()["si_loss"]
    value = world * B * a.steps / (ms / 1e3)
    e2e = world * B * a.steps / (ms_e2e / 1e3)
    # eager dispatch of the same step (what main.py's loop does call by call), for reference and for the launch count
    for _ in range(2):
        step(xd, td, True)
    n0 = depth_b200._lib.launch_count()
    ms_eager, _ = timed(min(a.steps, 3), False, eager_step)
    launches = (depth_b200._lib.launch_count() - n0) // min(a.steps, 3)
    ms_eager /= min(a.steps, 3)

    # ---- roofline of the tcgen05 conv kernels: CUDA events around every launch, on the launching stream --------
    hbm, tf_burst, tf_sus, src = peaks()
    roof = None
    rec = []
    orig_conv, orig_wg = ops._conv_tc_launch, ops._wgrad_tc

    def conv_hook(x, wp, Cout, KS, *rest, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = orig_conv(x, wp, Cout, KS, *rest, **kw); e.record()
        Bq, Hq, Wq, Cin = x.shape
        fuse = rest[8] if len(rest) > 8 else kw.get("fuse")
        has_res = len(rest) > 1 and (rest[1] is not None or rest[2] is not None)
        tag = ("conv+residual" if has_res else "conv") if fuse is None else \
              ("conv+bn_prologue" if fuse.pre_scale_shift else "conv+bn_backward")
        rec.append((tag, (Hq, Wq, Cin, Cout, KS), 2.0 * Bq * Hq * Wq * Cin * Cout * KS * KS, s, e))
        return r

    def wg_hook(x, g, Cin, Cout, KS, pre=None):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = orig_wg(x, g, Cin, Cout, KS, pre); e.record()
        Bq, Hq, Wq, _ = x.shape
        rec.append(("wgrad" if pre is None else "wgrad+bn_prologue", (Hq, Wq, Cin, Cout, KS),
                    2.0 * Bq * Hq * Wq * Cin * Cout * KS * KS, s, e))
        return r

    # every rank runs the instrumented steps (they contain the gradient all-reduce); rank 0 reports.  A device-side
    # sleep at the head of each step lets the (slower) eager host dispatch run ahead, so the CUDA events bracket GPU
    # execution only and not the host-side launch preparation of each call.
    ops._conv_tc_launch, ops._wgrad_tc = conv_hook, wg_hook
    nprof = 2
    for _ in range(nprof):
        torch.cuda._sleep(int(0.3 * 1.9e9))
        step(xd, td, True)
    torch.cuda.synchronize()
    ops._conv_tc_launch, ops._wgrad_tc = orig_conv, orig_wg
    if rank == 0:
        tot_ms = sum(s.elapsed_time(e) for _, _, _, s, e in rec)
        tot_fl = sum(f for _, _, f, _, _ in rec)
        per = {}
        for kind, shp, f, s, e in rec:
            k = (kind,) + shp
            d = per.setdefault(k, [0.0, 0.0, 0])
            d[0] += f; d[1] += s.elapsed_time(e); d[2] += 1
        if a.profile_layers:
            for k, d in sorted(per.items(), key=lambda kv: -kv[1][1]):
                sys.stderr.write(f"{str(k):58s} n={d[2]:3d} {d[1] / nprof:8.3f} ms/step {d[0] / d[1] / 1e9:8.1f} TFLOP/s\n")
        # dominant kernel = plain conv_tc_kernel launches on the shape that takes the most time in the step
        # (tensor-bound shapes only: both channel counts >= 64; the small-N full-resolution layers are HBM-bound and are
        # reported against the copy bandwidth below)
        plain = [(k, d) for k, d in per.items() if k[0] == "conv"]
        wide = [(k, d) for k, d in plain if min(k[3], k[4]) >= 64]
        top = max(wide or plain, key=lambda kv: kv[1][1])
        (_, Ht, Wt, cin_t, cout_t, ks_t), dt = top
        launch_ms = dt[1] / dt[2]
        ach = dt[0] / dt[1] / 1e9
        alg_bytes = 2.0 * B * Ht * Wt * (cin_t + cout_t)
        traffic, tsrc = (None, "no capture for this shape")
        if (Ht, Wt, cin_t, cout_t, ks_t, B) == (448, 576, 64, 64, 3, 32):
            # launch 4 of tools/ncu_kernels.py: conv 3x3 64->64 @448x576 (+stats), B = 32
            traffic, tsrc = _profile_traffic("ncu_kernels_r2.json", ["conv_tc.cu", "tc.cuh", "bn_fuse.cuh", "common.cuh"], index=4)
        roof = {"bound": "tensor",
                "kernel": f"conv_tc_kernel (tcgen05 implicit GEMM) {ks_t}x{ks_t} {cin_t}->{cout_t} @{Ht}x{Wt}, B={B}: the "
                          "shape with the largest share of the step (fusion-block convs and their data gradients)",
                "achieved": round(ach, 1), "peak": tf_sus, "unit": "TFLOP/s", "frac": round(ach / tf_sus, 4),
                "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step); the sustained figure was "
                               "taken power-capped - ncu's sm__pipe_tensor_cycles_active for this launch is in profiles/",
                "launch_ms": round(launch_ms, 4), "launches_per_step": dt[2] // nprof,
                "algorithmic_flops_per_launch": dt[0] / dt[2], "algorithmic_bytes_per_launch": alg_bytes,
                "traffic": traffic, "traffic_source": tsrc, "build_id": build,
                "all_tcgen05": {"kernels": "every conv_tc_kernel + wgrad_tc_kernel launch of the step (decoder, heads, "
                                           "cross-attention convs, EfficientNet 1x1s; incl. HBM-bound small-N layers)",
                                "achieved": round(tot_fl / (tot_ms / 1e3) / 1e12, 1), "unit": "TFLOP/s",
                                "frac": round(tot_fl / (tot_ms / 1e3) / 1e12 / tf_sus, 4),
                                "launches_per_step": len(rec) // nprof, "ms_per_step": round(tot_ms / nprof, 3),
                                "share_of_step": round((tot_ms / nprof) / (ms / a.steps), 3)}}
        fused = {}
        for k, d in per.items():
            if "+" in k[0]:
                e_ = fused.setdefault(k[0], [0.0, 0])
                e_[0] += d[1] / nprof; e_[1] += d[2] // nprof
        roof["bn_fused_launches"] = {k: {"ms_per_step": round(v[0], 3), "launches_per_step": v[1]} for k, v in fused.items()}
        # the full-resolution small-N convolutions are HBM-bound: report them against the copy bandwidth
        sm = [(k, d) for k, d in per.items() if k[0] == "conv" and k[1] * k[2] >= 448 * 576 and min(k[3], k[4]) <= 32]
        if sm:
            by = sum(2.0 * B * k[1] * k[2] * (k[3] + k[4]) * d[2] for k, d in sm)
            tms = sum(d[1] for _, d in sm)
            roof["hbm_bound_convs"] = {"kernels": "conv_tc_kernel, full-resolution layers with min(Cin,Cout) <= 32",
                                       "bound": "hbm", "achieved": round(by / (tms / 1e3) / 1e9, 1), "peak": hbm,
                                       "unit": "GB/s", "frac": round(by / (tms / 1e3) / 1e9 / hbm, 4),
                                       "ms_per_step": round(tms / nprof, 3)}

    # ---- the other BASELINE configs, measured by the driver at every N --------------------------------------------------
    ev = loop = c5 = None
    if not a.no_eval:
        ev = eval_kernel_leg(depth_b200, dev, rank, world, barrier, dist, hbm, build)
        loop = eval_loop_leg(depth_b200, model, dev, rank, world, barrier, dist)
    # release the train graph and its pools before the config-5 workload (the two do not coexist in a real job either)
    gstep.close()
    del gstep, red
    torch.cuda.empty_cache()
    if not a.no_config5:
        c5 = config5_leg(depth_b200, dev, rank, world, barrier, dist, tf_sus)

    cpu = tgb = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:      # reported at N = 1 only (the other ranks would idle on it)
        del model, opt
        torch.cuda.empty_cache()
        tgb = torch_gpu_baseline(B, 3)
        cpu = cpu_train_step(batch=4, steps=3, warmup=1)
        cpu["eval"] = cpu_eval_metrics()

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: default model (MidasNetSemantics f64, efficientnet_lite3 + dinov2_vits14 "
                                   "stand-ins, random init) bf16 training, batch 32/GPU synthetic RGB/depth 448x576, "
                                   "combined_loss weights 1/0/0/0, AdamW(1e-4,1e-4)",
                       "global_batch": world * B, "parallelism": f"dp{world}",
                       "l2_policy": "per-step working set (tens of GB of activations) >> 126 MB L2; no flush needed",
                       "encoders": ("efficientnet_lite3 stand-in trunk on libdepth_b200.so (depthwise + tcgen05 1x1 + fused BN); "
                                    if not a.torch_encoder else "efficientnet_lite3 stand-in run by PyTorch/cuDNN; ") +
                                   "frozen dinov2 stand-in run by PyTorch (bf16 autocast); decoder/fusion/heads/loss on "
                                   "libdepth_b200.so"},
            "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": int(xh.numel() * 4 + th.numel() * 4),
                    "d2h_bytes_per_step": 32, "ms_per_step": round(ms_e2e / a.steps, 3)},
            "gpu_launches": int(launches) * a.steps, "gpu_launches_per_step": int(launches),
            "execution": "whole train step captured once in a CUDA graph and replayed (depth_b200.GraphedTrainStep)",
            "eager_ms_per_step": round(ms_eager, 3), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "torch_gpu_baseline": tgb, "eval": ev, "eval_loop": loop, "config5": c5,
            "loss": last_loss, "build_id": build,
            "train_tflops_algorithmic": round(3 * FWD_GFLOP_PER_IMG * value / 1e3, 1),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Orderly teardown: the captured graph (which recorded NCCL collectives) is already released above; synchronise,
        # meet at a barrier, destroy the process group.  Round 1 left through os._exit because destroying the group with
        # the graph alive hung; a watchdog keeps that exit as the last resort so a teardown problem can never hang the
        # driver's scaling run (the JSON line is already out).
        sys.stdout.flush()
        sys.stderr.flush()
        wd = threading.Timer(30.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        wd.cancel()


def cpu_eval_metrics():
    """the reference's evaluation reductions (evaluation.py:157-166 call pattern: SI-RMSE + AbsRel + 3 delta) on the host
    cores, through the oracle port: CPU Gpx/s next to the GPU kernel's"""
    import torch
    from oracle import losses as ol
    torch.set_num_threads(os.cpu_count())
    EB = 32
    g = torch.Generator().manual_seed(7)
    t = torch.rand(EB, 1, H, W, generator=g) * 9.9 + 0.1
    p = t * torch.exp(0.1 * torch.randn(EB, 1, H, W, generator=g)) * 1.3

    def once():
        ol.scale_invariant_loss(p, t, sqroot=True).item()
        ol.absolute_relative_error(p, t).item()
        for j in (1, 2, 3):
            ol.delta_thres(p, t, 1.05 ** j).item()

    once()
    dts = []
    for _ in range(3):
        t0 = time.perf_counter()
        once()
        dts.append(time.perf_counter() - t0)
    sec = sum(dts) / len(dts)
    return {"value": round(EB * H * W / sec / 1e9, 4), "unit": "Gpx/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"3 timed calls of the five metric functions on batch {EB} at 448x576 after 1 warm-up; fp32"}


def cpu_train_step(batch, steps, warmup):
    """the reference's CPU path (oracle port, pinned to the reference by oracle/make_golden.py) on the host cores"""
    import torch
    from oracle import cases, fixtures as fx, losses as ol
    standins = fx.load_standins()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = cases.build_oracle_semantics(standins).train()
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
    x, t = synthetic_batch(batch, 1234)
    cfg = fx.loss_config()
    dts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        out = model(x).unsqueeze(1)
        loss, _ = ol.combined_loss(out, t, cfg, rgb=x)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            dts.append(dt)
    sec = sum(dts) / len(dts)
    return {"value": round(batch / sec, 4), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{steps} train step(s) of batch {batch} at 448x576 after {warmup} warm-up (configs[0]); fp32, "
                      f"torch {torch.__version__} CPU, {os.cpu_count()} threads", "sec_per_step": round(sec, 3)}


def run_reference(a):
    """--impl reference: the reference's own CPU implementation (Python: the oracle port) on this box's host cores, on
    OUR arm's configuration family: the default model, batch 4 per step (configs[0]; the reference's own batch size,
    config.yaml:15).  The batch is never reduced; when K + W steps of ~4 s would not end within a few minutes the number
    of timed steps is cut instead and the line says how many were run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget_steps = 40                                   # ~4 s per batch-4 step on 16 cores -> <= ~3 minutes
    warm = max(1, min(a.warmup, 2))
    steps = max(3, min(a.steps, budget_steps - warm))
    r = cpu_train_step(batch=4, steps=steps, warmup=warm)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": a.gpus,
            "steps": steps, "warmup": warm, "steps_requested": a.steps, "warmup_requested": a.warmup,
            "ms_per_step": round(r["sec_per_step"] * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[0]: default model (MidasNetSemantics f64, stand-in encoders, random init), one train "
                                   "step = fwd + combined_loss 1/0/0/0 + bwd + AdamW, batch 4 synthetic RGB/depth 448x576, on "
                                   "the host CPU (bounded sample of the GPU arm's batch-32 workload: same model, same step)",
                       "global_batch": 4, "parallelism": "cpu"},
            "cpu_baseline": {"kind": r["kind"], "cores": r["cores"], "sample": r["sample"], "value": r["value"],
                             "unit": "images/s"},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
