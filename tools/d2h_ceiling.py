#!/usr/bin/env python3
"""What the box's host fabric gives to device->host copies: plain page-locked cudaMemcpyAsync of one 4K frame (99.5 MB) per
GPU, first on rank 0 alone, then on all ranks at once - the ceiling bench.py's end-to-end figures are bound by. Run under
torchrun (one rank per GPU) or alone. One JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 tools/d2h_ceiling.py
"""
import json, os, time
import torch

rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
n = 3840 * 2160 * 3  # u32 words of one frame's three planes
dev = torch.empty(n, dtype=torch.int32, device=f"cuda:{local_rank}")
pin = torch.empty(n, dtype=torch.int32).pin_memory()
steps = 30


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()


def run(active: bool) -> float:
    for _ in range(3):
        if active:
            pin.copy_(dev, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    if active:
        for _ in range(steps):
            pin.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    return (t1 - t0) if active else 0.0


alone = run(rank == 0)
together = run(True)
out = {"world": world, "bytes_per_copy": n * 4}
if dist is not None:
    t = torch.tensor([together], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    together = float(t.item())
if rank == 0:
    out["one_gpu_gbs"] = n * 4 * steps / alone / 1e9
    out["all_gpus_aggregate_gbs"] = world * n * 4 * steps / together / 1e9
    out["all_gpus_per_gpu_gbs"] = n * 4 * steps / together / 1e9
    print(json.dumps(out))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
