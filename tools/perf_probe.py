#!/usr/bin/env python3
"""Quick kernel-time probe over several scenes (GPU box). Prints ms/frame (L2 flushed) and Grays/s per scene.

    python tools/perf_probe.py [--mips VIEWING_DISTANCE] [scene ...]
--mips: switch the trees' MIP maps on (default strategy) and render through get_by_ray_at_lod at that viewing distance
(the LOD kernel variant); "frustum" uses each camera's viewport.frustum.z like the reference's shader does."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import shocovox_b200 as S  # noqa: E402
from shocovox_b200 import scenes  # noqa: E402

CASES = {
    "dot_cube_1080p": (lambda: scenes.dot_cube_scene(), lambda: scenes.dot_cube_camera(zoom=True), (1920, 1080)),
    "dot_cube_1080p_fov": (lambda: scenes.dot_cube_scene(), lambda: scenes.dot_cube_camera(zoom=False), (1920, 1080)),
    "dot_cube_4k": (lambda: scenes.dot_cube_scene(), lambda: scenes.dot_cube_camera(zoom=True), (3840, 2160)),
    "cpu_render_1080p": (lambda: scenes.cpu_render_scene(), lambda: scenes.cpu_render_camera(), (1920, 1080)),
    "cpu_render_4k": (lambda: scenes.cpu_render_scene(), lambda: scenes.cpu_render_camera(), (3840, 2160)),
    "colonnade_4k": (lambda: scenes.colonnade_scene(), lambda: scenes.colonnade_camera(), (3840, 2160)),
    "terrain_512_8_4k": (lambda: scenes.terrain_scene(512, 8, 4321, 1, shell=4), lambda: scenes.terrain_camera(512), (3840, 2160)),
    "terrain_1024_8_1080p": (lambda: scenes.terrain_scene(1024, 8, 4321, 1, shell=4), lambda: scenes.orbit_cameras(1024, 256)[0], (1920, 1080)),
    "sponza_2048_32_4k": (lambda: scenes.colonnade_scene(2048, 32), lambda: scenes.colonnade_camera(2048), (3840, 2160)),
    "minecraft_1024_32_4k": (lambda: scenes.terrain_scene(1024, 32, 1234, 4, shell=8, name="minecraft"), lambda: scenes.terrain_camera(1024), (3840, 2160)),
    "minecraft_256_32_4k": (lambda: scenes.terrain_scene(256, 32, 1234, 4, shell=8, name="minecraft"), lambda: scenes.terrain_camera(256), (3840, 2160)),
}


def main():
    argv = sys.argv[1:]
    mips = None
    if "--mips" in argv:
        i = argv.index("--mips")
        mips = argv[i + 1]
        del argv[i:i + 2]
    names = argv or list(CASES)
    trees = {}
    out = {}
    for name in names:
        mk_scene, mk_cam, res = CASES[name]
        sc = mk_scene()
        if sc.name not in trees:
            t0 = time.time()
            tree = scenes.build_tree(sc, S.Octree)
            tb = time.time() - t0
            if mips is not None:
                t0 = time.time()
                tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
                print(f"  MIP maps of {sc.name} built in {time.time() - t0:.1f} s", flush=True)
            trees[sc.name] = (tree, tb)
        tree, tb = trees[sc.name]
        cam = mk_cam()
        host = S.OctreeGPUHost(tree)
        view = host.create_new_view(64, S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov), res)
        if cam.glass_at_frustum_z:
            view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
        if mips is not None:
            view.set_viewing_distance(float(cam.frustum[2]) if mips == "frustum" else float(mips))
        ms = []
        for i in range(12):
            view.flush_l2()
            k = view.render(sync=True)["kernel_ms"]
            if i >= 2:
                ms.append(k)
        warm = []
        for i in range(8):
            k = view.render(sync=True)["kernel_ms"]
            if i >= 2:
                warm.append(k)
        f = view.render_to_host()
        hits = int((f["hit_id"] != S.MISS).sum())
        st = host.stats()
        out[name] = {"ms_warm": float(np.mean(warm)), "ms": float(np.mean(ms)), "ms_min": float(np.min(ms)), "grays": res[0] * res[1] / np.mean(ms) / 1e6,
                     "hits": hits, "nodes": st["nodes"], "bricks": st["bricks"], "depth": st["depth"], "MB": st["total_bytes"] / 1e6,
                     "build_s": round(tb, 2)}
        print(f"{name:24s} {out[name]['ms']:9.4f} ms (warm L2 {out[name]['ms_warm']:8.4f})  {out[name]['grays']:8.2f} Grays/s  hits {hits:8d}  nodes {st['nodes']:6d} "
              f"bricks {st['bricks']:6d} depth {st['depth']} tree {st['total_bytes'] / 1e6:7.1f} MB", flush=True)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / ("perf_probe.json" if mips is None else f"perf_probe_mips_{mips}.json")).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
