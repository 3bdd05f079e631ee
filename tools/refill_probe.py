#!/usr/bin/env python3
"""Lane refill A/B: the static schedule vs persistent warps pulling tiles vs persistent warps whose
finished lanes take new pixels while the others keep their place in the tree (svx_view_set_schedule 0 / 1 / 2), on whole
frames and on rank 0's share of a frame split over 8 GPUs. The refill schedule's three parameters (node-loop iterations per
round, idle lanes before a hand-out, 8x4 tiles per ticket) are swept. One GPU; ms per frame, L2 flushed; every frame compared byte for byte with
the static schedule's.

    python tools/refill_probe.py [workload ...]
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
SWEEP = [(64, 16, 1), (24, 8, 8), (64, 16, 8), (24, 8, 1), (24, 16, 1), (64, 8, 1), (64, 24, 1), (64, 16, 2), (256, 16, 1), (256, 32, 1), (8, 16, 1)]
out = {}
for name in names:
    scene, cams, res, _ = bench.make_workload(name)
    cam = cams[0]
    tree = scenes.build_tree(scene, S.Octree)
    host = S.OctreeGPUHost(tree, 0)
    rec = {}
    for world in (1, 8):
        modes = [("static", 0, None), ("persistent", 1, None)] + [(f"refill_s{s}_i{i}_u{u}", 2, (s, i, u)) for s, i, u in (SWEEP if world == 1 else SWEEP[:1])]
        reference_frame = None
        equal = True
        for label, schedule, params in modes:
            if params:
                os.environ["SVX_REFILL_STEPS"], os.environ["SVX_REFILL_MIN_IDLE"], os.environ["SVX_REFILL_UNIT"] = (str(p) for p in params)
            view = host.create_new_view(64, S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov), res)
            if cam.glass_at_frustum_z:
                view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
            view.set_schedule(schedule)
            if world > 1:
                view.set_shard(0, world, 8)
            ms = []
            for i in range(16):
                view.flush_l2()
                k = view.render(sync=True)["kernel_ms"]
                if i >= 4:
                    ms.append(k)
            rec[f"world{world}_{label}"] = round(float(np.mean(ms)), 4)
            frame = view.render_to_host()
            rows = np.array([r for r in range(res[1]) if (r // 8) % world == 0])
            planes = [frame[k][rows].view(np.uint32) for k in ("hit_id", "albedo", "distance")]
            if reference_frame is None:
                reference_frame = planes
            else:
                equal &= all(bool(np.array_equal(a, b)) for a, b in zip(reference_frame, planes))
            del view
        rec[f"world{world}_frames_equal"] = equal
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
print(json.dumps(out))
