#!/usr/bin/env python3
"""End-to-end read-back probe on one GPU: the box's device->host ceiling (plain pinned copies) against the pipelined and the
synchronised svx_view_render_to_host paths, for every subset of the three planes. Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import shocovox_b200 as S
from shocovox_b200 import scenes

name = sys.argv[1] if len(sys.argv) > 1 else "sponza_4k"
steps = 30
scene, cams, res, _ = bench.make_workload(name)
w, h = res
n_px = w * h
tree = scenes.build_tree(scene, S.Octree)
host = S.OctreeGPUHost(tree, 0)
cam = cams[0]
view = host.create_new_view(64, S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov), res)
if cam.glass_at_frustum_z:
    view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
out = {"workload": name, "n_px": n_px}
# ceiling: torch pinned copies of the same sizes
dev = torch.empty(3 * n_px, dtype=torch.int32, device="cuda:0")
pin = torch.empty(3 * n_px, dtype=torch.int32).pin_memory()
for label, n in (("d2h_1_plane", n_px), ("d2h_3_planes_one_copy", 3 * n_px)):
    for _ in range(3):
        pin[:n].copy_(dev[:n], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        pin[:n].copy_(dev[:n], non_blocking=True)
    torch.cuda.synchronize()
    out[label + "_gbs"] = n * 4 * steps / (time.perf_counter() - t0) / 1e9
sets = [[torch.empty(n_px, dtype=torch.int32).pin_memory() for _ in range(3)] for _ in range(2)]
ptrs = [[t.data_ptr() for t in s] for s in sets]
k = view.render(sync=True)["kernel_ms"]
out["kernel_ms"] = k
for label, mask in (("hit", (1, 0, 0)), ("hit_alb", (1, 1, 0)), ("hit_dist", (1, 0, 1)), ("alb_dist", (0, 1, 1)), ("all", (1, 1, 1))):
    sel = lambda p: [q if m else 0 for q, m in zip(p, mask)]
    nbytes = sum(mask) * n_px * 4
    for i in range(3):
        view.render_to_host_async_ptr(*sel(ptrs[i & 1])); view.wait_host(1)
    view.wait_host(0)
    t0 = time.perf_counter()
    for i in range(steps):
        view.render_to_host_async_ptr(*sel(ptrs[i & 1])); view.wait_host(1)
    view.wait_host(0)
    t = (time.perf_counter() - t0) / steps
    t0 = time.perf_counter()
    for i in range(steps):
        view.render_to_host_ptr(*sel(ptrs[0]))
    ts = (time.perf_counter() - t0) / steps
    out[label] = {"pipelined_ms": t * 1e3, "pipelined_gbs": nbytes / t / 1e9, "sync_ms": ts * 1e3, "sync_copy_gbs": nbytes / max(ts - k * 1e-3, 1e-9) / 1e9}
print(json.dumps(out))
