#!/usr/bin/env python3
"""Where does the time of a tile-sharded frame go? Run under torchrun (one rank per GPU). For the workload's pose 0:
the whole frame on one GPU, this rank's shard rendered locally (no exchange), and the fused gather under every
SVX_GATHER_TUNING setting and both wire formats - per-rank device times (rank 0: viewport kernel + wait for the slowest
peer = the frame; peers: their own viewport kernel). One JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 tools/gather_probe.py [workload] [steps]
"""
import json, os, sys
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import shocovox_b200 as S  # noqa: E402
from shocovox_b200 import distributed as D, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sponza_4k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bands = [int(b) for b in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["8"])]
TUNINGS = [int(t) for t in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["0", "4", "16"])]
rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
scene, cams, res, _ = bench.make_workload(name)
cam = cams[0]
tree = scenes.build_tree(scene, S.Octree)
host = S.OctreeGPUHost(tree, local_rank)
vp = S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov)


def new_view():
    v = host.create_new_view(64, vp, res)
    if cam.glass_at_frustum_z:
        v.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
    return v


def gathered(times):
    t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=f"cuda:{local_rank}")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [round(float(o.item()), 4) for o in out]


def timed(view):
    for _ in range(3):
        view.flush_l2()
        view.render(sync=True)
    view.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    ms = []
    for _ in range(steps):
        view.flush_l2()
        ms.append(view.render(sync=True)["kernel_ms"])
    view.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    return gathered(ms)


report = {"workload": name, "world": world, "steps": steps}
whole = new_view()
report["whole_frame_ms_per_rank"] = timed(whole)
ref = whole.render_to_host() if rank == 0 else None
for band in bands:
    shard = new_view()
    shard.set_shard(rank, world, band)
    report[f"local_shard_band{band}_ms_per_rank"] = timed(shard)
    shard.set_schedule(True)
    report[f"local_shard_band{band}_persistent_ms_per_rank"] = timed(shard)
    del shard
    for wire, wname in ((S.WIRE_THREE_PLANES, "12B"), (S.WIRE_ID_DISTANCE, "8B")):
        for tuning in TUNINGS:
            os.environ["SVX_GATHER_TUNING"] = str(tuning)
            v = new_view()
            D.open_gather(v, rank, world, band, wire)
            key = f"gather_band{band}_{wname}_tuning{tuning}"
            report[key + "_ms_per_rank"] = timed(v)
            if not tuning & 4:  # the frame is only right when the peers store into the root
                v.render(sync=False)
                v.synchronize()
                dist.barrier()
                if rank == 0:
                    got = v.read_frame()
                    report[key + "_equal"] = all(bool(np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32))) for k in ("hit_id", "albedo", "distance"))
            dist.barrier()
            if rank != 0:
                v.gather_close()
            dist.barrier()
            if rank == 0:
                v.gather_close()
            del v
os.environ.pop("SVX_GATHER_TUNING", None)
if rank == 0:
    print(json.dumps(report))
dist.barrier()
dist.destroy_process_group()
