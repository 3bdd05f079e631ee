#!/usr/bin/env python3
"""Per-pixel timing of one frame (or of one rank's shard of it) with the MEASUREMENT build of the library
(-DSVX_PIXEL_TIMING=1, built here as shocovox_b200/libpixeltiming.so): how long do the slowest rays run, where are they on
the screen, and how long is the tail of the launch during which most of the machine idles? That tail does not shrink when
the frame is split over more GPUs - it is what bounds strong scaling. One JSON line.

    SVX_LIB=shocovox_b200/libpixeltiming.so python tools/pixel_timing.py [workload] [world]
"""
import json, os, sys
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import shocovox_b200 as S  # noqa: E402
from shocovox_b200 import scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sponza_4k"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
scene, cams, res, _ = bench.make_workload(name)
w, h = res
cam = cams[0]
tree = scenes.build_tree(scene, S.Octree)
host = S.OctreeGPUHost(tree, 0)
out = {"workload": name, "library": os.environ.get("SVX_LIB", "default")}
for shard_world in (1, world):
    for persistent in (False, True):
        view = host.create_new_view(64, S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov), res)
        if cam.glass_at_frustum_z:
            view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
        if shard_world > 1:
            view.set_shard(0, shard_world, 8)
        view.set_schedule(persistent)
        for _ in range(3):
            view.flush_l2()
            k = view.render(sync=True)["kernel_ms"]
        view.flush_l2()
        f = view.render_to_host()
        rows = np.array([r for r in range(h) if (r // 8) % shard_world == 0])
        cyc = f["hit_id"][rows].astype(np.float64)
        t0 = f["albedo"][rows].astype(np.int64)
        t1 = f["distance"][rows].view(np.uint32).astype(np.int64)
        start = t0.min()
        rel0, rel1 = (t0 - start) & 0xFFFFFFFF, (t1 - start) & 0xFFFFFFFF
        end = rel1.max()
        dur_ns = (rel1 - rel0).astype(np.float64)
        # machine occupancy over time: pixels in flight per microsecond bin
        bins = np.arange(0, end + 1000, 1000)
        inflight = np.zeros(len(bins))
        np.add.at(inflight, np.clip(rel0.ravel() // 1000, 0, len(bins) - 1), 1)
        np.add.at(inflight, np.clip(rel1.ravel() // 1000 + 1, 0, len(bins) - 1), -1)
        inflight = np.cumsum(inflight)
        full = inflight.max()
        tail_start = np.argmax(inflight[::-1] > 0.5 * full)
        key = f"world{shard_world}_{'persistent' if persistent else 'static'}"
        # the slowest pixels, by 8x4 tile position
        flat = np.argsort(dur_ns.ravel())[-5:]
        slow = [(int(rows[i // w]), int(i % w), float(dur_ns.ravel()[i] / 1000)) for i in flat]
        out[key] = {"kernel_ms": k, "span_us": float(end / 1000), "pixel_us_mean": float(dur_ns.mean() / 1000), "pixel_us_p50": float(np.percentile(dur_ns, 50) / 1000),
                    "pixel_us_p99": float(np.percentile(dur_ns, 99) / 1000), "pixel_us_p999": float(np.percentile(dur_ns, 99.9) / 1000), "pixel_us_max": float(dur_ns.max() / 1000),
                    "cycles_mean": float(cyc.mean()), "cycles_max": float(cyc.max()),
                    "us_below_half_occupancy_at_the_end": float(tail_start), "peak_pixels_in_flight": float(full),
                    "row_mean_us_by_image_eighth": [float(dur_ns[(rows >= a) & (rows < a + h // 8)].mean() / 1000) for a in range(0, h, h // 8)][:8],
                    "slowest_pixels_row_x_us": slow}
        del view
print(json.dumps(out))
