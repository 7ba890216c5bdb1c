"""H2D bandwidth from pinned memory with all ranks copying at once (torchrun)."""
import os, torch, torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = torch.empty(132 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    d.copy_(h, non_blocking=True)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 20
print(f"rank {rank}: H2D 132 MiB in {ms:.2f} ms = {132 * 1.048576 / ms:.1f} GB/s", flush=True)
if world > 1:
    dist.barrier()
    os._exit(0)
