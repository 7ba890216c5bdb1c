"""Drop-in for the metric loop of the reference's ``src/evaluation.py`` (evaluation.py:138-186).

The reference walks a DataLoader, calls scale_invariant_loss(sqroot=True), absolute_relative_error and three
delta_thres per batch (five passes over pred/target plus six `.item()`-style syncs), weights each by the batch length,
clips to N_SAMPLES and prints the averages.  Here every batch costs ONE fused kernel launch
(`util.evaluation_metrics`: the streaming kernel, each input byte read from HBM once), the weighted sums stay on the device, and with
`torch.distributed` initialised the samples are sharded by rank and the partial sums meet in one all-reduce of
2 + N_DELTA + 1 doubles.
"""
import torch

from . import distributed as D
from . import util

BASE_THRES = 1.05        # evaluation.py:27
N_DELTA = 3              # evaluation.py:28


def evaluate_batches(model, batches, device, n_samples=None, base_thres=BASE_THRES, n_delta=N_DELTA, fast_math=None,
                     local_shard=False):
    """`batches` yields (rgb, depth_gt, ...) like the reference's DataLoader (evaluation.py:143).  Returns a dict with the
    reference's three printed quantities: 'si_rmse', 'abs_rel', 'delta' (list of n_delta) and 'samples'.

    Every rank walks the whole iterable and keeps the batches whose index is congruent to its rank (shard by batch);
    n_samples clips the total exactly as evaluation.py:169-177 does (a partially used last batch is weighted by the
    number of samples taken from it, with the batch-level metric values - the reference's arithmetic).
    local_shard=True: `batches` already holds only this rank's samples (a DistributedSampler-style loader): nothing is
    skipped, and n_samples clips the local stream."""
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    thr = [base_thres ** j for j in range(1, n_delta + 1)]
    sums = torch.zeros(2 + n_delta, dtype=torch.float64, device=device)
    count = 0
    seen = 0
    model.eval()
    with torch.no_grad():
        for idx, batch in enumerate(batches):
            rgb, gt = batch[0], batch[1]
            b = rgb.shape[0]
            take = b if n_samples is None else max(0, min(b, n_samples - seen))
            seen += b
            if take == 0:
                break
            if not local_shard and idx % world != rank:
                continue
            pred = model(rgb.to(device, non_blocking=True))
            if pred.dim() == 3:
                pred = pred.unsqueeze(1)
            m = util.evaluation_metrics(pred, gt.to(device, non_blocking=True), thresholds=thr, fast_math=fast_math)
            sums += m.double() * take            # batch means weighted by the samples counted (evaluation.py:158-166)
            count += take
    tot = D.all_reduce_metric_sums(sums.tolist() + [float(count)], device=device if world > 1 else None)
    n = max(tot[-1], 1.0)
    return {"si_rmse": tot[0] / n, "abs_rel": tot[1] / n, "delta": [v / n for v in tot[2:2 + n_delta]], "samples": int(tot[-1])}


def evaluate_model(model, val_loader, device):
    """Drop-in for the reference's ``main.evaluate_model`` (main.py:254-392): eval-mode forward, bilinear resize of the
    prediction to the target size (align_corners=True), then MAE / RMSE / siRMSE / REL / unaligned delta < 1.25^k over the
    validation set, normalised by N * C * H * W exactly as the reference does.  Each batch costs one fused moments pass
    and one counting pass (util.evaluate_model_sums) instead of nine tensor passes and a per-image numpy loop; with
    torch.distributed initialised the batches are sharded by rank and the seven sums meet in one all-reduce."""
    import math
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    model.eval()
    keys = ("abs", "sq", "rel", "sirmse", "d1", "d2", "d3")
    acc = dict.fromkeys(keys, 0.0)
    total_samples, total_pixels = 0, None
    with torch.no_grad():
        for idx, (inputs, targets, _filenames) in enumerate(val_loader):
            if total_pixels is None:
                total_pixels = targets.shape[1] * targets.shape[2] * targets.shape[3]
            if idx % world != rank:
                continue
            inputs, targets = inputs.to(device, non_blocking=True), targets.to(device, non_blocking=True)
            total_samples += inputs.size(0)
            outputs = model(inputs).unsqueeze(1)
            sums = util.evaluate_model_sums(outputs, targets)
            for k in keys:
                acc[k] += float(sums[k])
    tot = D.all_reduce_metric_sums([acc[k] for k in keys] + [float(total_samples)], device=device if world > 1 else None)
    n_s = max(tot[-1], 1.0)
    n = n_s * (total_pixels or 1)
    return {"MAE": tot[0] / n, "RMSE": math.sqrt(tot[1] / n), "siRMSE": tot[3] / n_s, "REL": tot[2] / n,
            "Delta1": tot[4] / n, "Delta2": tot[5] / n, "Delta3": tot[6] / n}
