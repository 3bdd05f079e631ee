#!/usr/bin/env python3
"""Run under torchrun (one rank per GPU): ONE frame split into row bands over the ranks must equal the 1-GPU frame byte
for byte, for
  (a) the fused gather (csrc/multi_gpu.cu): every rank's viewport kernel stores straight into rank 0's framebuffer
      through a CUDA-IPC mapping, go / done flags on the devices, no host barrier per frame - both wire formats;
  (b) the comparison path: compact bands + one NCCL all-gather per plane + de-interleave (shocovox_b200.distributed).
Also times both (device events on rank 0 for (a): viewport kernel + wait for the slowest peer).

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
    ap.add_argument("--skip-nccl", action="store_true")
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

    def same(got, ref):
        return all(bool(np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32))) for k in ("hit_id", "albedo", "distance"))

    report = {"world": world, "workload": args.workload, "resolution": list(res)}
    # reference frame: every rank renders the full frame alone (also the 1-GPU timing)
    full = new_view()
    ref = full.render_to_host()
    ms = []
    for _ in range(args.steps):
        full.flush_l2()
        ms.append(full.render(sync=True)["kernel_ms"])
    report["single_gpu_ms"] = float(np.mean(ms))
    ok_all = True

    # (a) fused gather, both wire formats
    for wire, name in ((S.WIRE_THREE_PLANES, "fused_12B"), (S.WIRE_ID_DISTANCE, "fused_8B")):
        v = new_view()
        D.open_gather(v, rank, world, args.band, wire)
        ok = True
        for i in range(3):  # several frames: the sequence numbers advance on every member
            v.render(sync=False)
            v.synchronize()
            if rank == 0:
                ok &= same(v.read_frame(), ref)
        dist.barrier()
        times = []
        for i in range(args.steps):
            v.flush_l2()
            if rank == 0:
                times.append(v.render(sync=True)["kernel_ms"])
            else:
                v.render(sync=False)
        v.synchronize()
        dist.barrier()
        if rank == 0:
            ok &= same(v.read_frame(), ref)
        flag = torch.tensor([int(ok)], device=f"cuda:{local_rank}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_all &= bool(flag.item())
        report[name] = {"equal_to_single_gpu": bool(flag.item()), "ms_per_frame_device": float(np.mean(times)) if times else None}
        dist.barrier()
        if rank != 0:
            v.gather_close()
        dist.barrier()
        if rank == 0:
            v.gather_close()

    # (b) compact bands + NCCL all-gather
    if not args.skip_nccl:
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
        ok_a = all(bool(np.array_equal(got.cpu().numpy().view(np.uint32), ref[name].view(np.uint32)))
                   for got, name in zip(out, ("hit_id", "albedo", "distance")))
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_a()
        torch.cuda.synchronize()
        dist.barrier()
        flag = torch.tensor([int(ok_a)], device=f"cuda:{local_rank}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_all &= bool(flag.item())
        report["nccl_gather"] = {"equal_to_single_gpu": bool(flag.item()), "ms_per_frame_wall": (time.perf_counter() - t0) / args.steps * 1e3}
    if rank == 0:
        report["all_ranks_ok"] = ok_all
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
