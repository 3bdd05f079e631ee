#!/usr/bin/env python3
"""Static schedule in raster order vs the static schedule dispatching its 16x8-pixel blocks heaviest-first (by their cost in
the previous frame; SVX_CTA_ORDER) vs persistent warps pulling tiles in raster order, on whole frames and on rank 0's share of a frame split over 2 / 4 / 8 GPUs (the strong-scaling
case: the tail of the launch does not shrink with the share). One GPU; ms per frame, L2 flushed; frames compared byte for byte.

    python tools/schedule_probe.py [workload ...]
"""
import json, os, sys
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import shocovox_b200 as S  # noqa: E402
from shocovox_b200 import scenes  # noqa: E402

names = sys.argv[1:] or ["sponza_4k", "minecraft_4k", "terrain_poses_1080p", "dot_cube_1080p"]
out = {}
for name in names:
    scene, cams, res, _ = bench.make_workload(name)
    cam = cams[0]
    tree = scenes.build_tree(scene, S.Octree)
    host = S.OctreeGPUHost(tree, 0)
    rec = {}
    for world in (1, 2, 4, 8):
        frames = {}
        for mode in ("static", "static_ordered_12", "static_ordered_30", "static_ordered_60", "static_ordered_100"):
            os.environ["SVX_CTA_ORDER"] = "2" if mode.startswith("static_ordered") else "0"
            os.environ["SVX_CTA_ORDER_HEAD_PCT"] = mode.rsplit("_", 1)[1] if mode.startswith("static_ordered") else "12"
            view = host.create_new_view(64, S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov), res)
            if cam.glass_at_frustum_z:
                view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
            if world > 1:
                view.set_shard(0, world, 8)
            ms = []
            for i in range(20):
                view.flush_l2()
                k = view.render(sync=True)["kernel_ms"]
                if i >= 5:
                    ms.append(k)
            rec[f"world{world}_{mode}"] = round(float(np.mean(ms)), 4)
            frames[mode] = view.render_to_host()
            del view
        rows = np.array([r for r in range(res[1]) if (r // 8) % world == 0])
        rec[f"world{world}_frames_equal"] = all(
            bool(np.array_equal(frames["static"][k][rows].view(np.uint32), frames[m][k][rows].view(np.uint32)))
            for m in ("static_ordered_12", "static_ordered_30", "static_ordered_60", "static_ordered_100") for k in ("hit_id", "albedo", "distance"))
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
os.environ.pop("SVX_CTA_ORDER", None)
os.environ.pop("SVX_CTA_ORDER_HEAD_PCT", None)
print(json.dumps(out))
