#!/usr/bin/env python3
"""Run under torchrun (one rank per GPU): one frame split into row bands over the ranks must equal the 1-GPU frame
byte for byte, for (a) the NCCL all-gather of compact bands and (b) the fused variant where every rank's traversal
kernel stores straight into rank 0's framebuffer through CUDA-IPC peer mappings. Also times both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--workload cpu_render_1080p] [--steps 20]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import shocovox_b200 as S  # noqa: E402
from shocovox_b200 import distributed as D, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cpu_render_1080p")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--band", type=int, default=8)
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scene, cams, res, _ = bench.make_workload(args.workload)
    cam = cams[0]
    w, h = res
    tree = scenes.build_tree(scene, S.Octree)
    host = S.OctreeGPUHost(tree, local_rank)
    vp = S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov)

    def new_view():
        v = host.create_new_view(64, vp, res)
        if cam.glass_at_frustum_z:
            v.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
        return v

    report = {"world": world, "workload": args.workload, "resolution": list(res)}
    # reference frame: every rank renders the full frame alone (also the 1-GPU timing)
    full = new_view()
    ref = full.render_to_host()
    ms = []
    for _ in range(args.steps):
        full.flush_l2()
        ms.append(full.render(sync=True)["kernel_ms"])
    report["single_gpu_ms"] = float(np.mean(ms))

    # (a) compact bands + NCCL all-gather
    lr = D.padded_local_rows(h, world, args.band)
    va = new_view()  # framebuffer is (h, w); only the first `lr` rows are used when compact
    va.set_shard(rank, world, args.band)
    va.set_compact_rows(True)
    ptrs = va.frame_pointers()
    planes = [D.device_tensor(p, (h, w), dt, local_rank)[:lr] for p, dt in zip(ptrs, ("<i4", "<i4", "<f4"))]
    stream = torch.cuda.ExternalStream(va.cuda_stream(), device=local_rank)

    def step_a():
        va.render(sync=False)
        with torch.cuda.stream(stream):
            return [D.gather_bands(p, h, world, args.band) for p in planes]

    out = step_a()
    torch.cuda.synchronize()
    ok_a = True
    for got, name in zip(out, ("hit_id", "albedo", "distance")):
        ok_a &= bool(np.array_equal(got.cpu().numpy().view(np.uint32), ref[name].view(np.uint32)))
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_a()
    torch.cuda.synchronize()
    dist.barrier()
    t_a = (time.perf_counter() - t0) / args.steps * 1e3
    report["nccl_gather"] = {"equal_to_single_gpu": ok_a, "ms_per_frame_wall": t_a}

    # (b) fused: kernels store straight into rank 0's framebuffer over NVLink (CUDA IPC peer mapping)
    target = new_view()  # rank 0's is the destination
    blob = D.exchange_ipc_handles(target.export_frame_ipc(), src_rank=0)
    vb = new_view()
    vb.set_shard(rank, world, args.band)
    if rank != 0:
        vb.set_peer_frame_ipc(blob)
    else:
        # rank 0 renders its own bands into the same destination buffers
        vb = target
        vb.set_shard(0, world, args.band)

    def step_b():
        vb.render(sync=False)
        vb.synchronize()

    step_b()
    dist.barrier()
    ok_b = True
    if rank == 0:
        p = target.frame_pointers()
        for ptr, dt, name in zip(p, ("<i4", "<i4", "<f4"), ("hit_id", "albedo", "distance")):
            got = D.device_tensor(ptr, (h, w), dt, local_rank).cpu().numpy()
            ok_b &= bool(np.array_equal(got.view(np.uint32), ref[name].view(np.uint32)))
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_b()
        dist.barrier()
    torch.cuda.synchronize()
    t_b = (time.perf_counter() - t0) / args.steps * 1e3
    report["fused_peer_stores"] = {"equal_to_single_gpu": ok_b if rank == 0 else None, "ms_per_frame_wall": t_b}
    flags = torch.tensor([int(ok_a), int(ok_b)], device=f"cuda:{local_rank}")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank != 0:
        vb.set_peer_frame_ipc(None)
    dist.barrier()
    if rank == 0:
        report["all_ranks_ok"] = bool(flags.min().item() == 1)
        print(json.dumps(report))
    dist.destroy_process_group()
    return 0 if flags.min().item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
