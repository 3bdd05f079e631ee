#!/usr/bin/env python3
"""Render-data upload cost (GPU box): full upload vs incremental svx_gpu_host_reload after small edits.

usage: tools/reload_probe.py [minecraft|terrain|dot_cube]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import shocovox_b200 as S  # noqa: E402
from shocovox_b200 import scenes  # noqa: E402

SCENES = {
    "minecraft": lambda: scenes.terrain_scene(1024, 32, 1234, 4, shell=8, name="minecraft"),
    "terrain": lambda: scenes.terrain_scene(1024, 8, 4321, 1, shell=4),
    "dot_cube": lambda: scenes.dot_cube_scene(),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "dot_cube"
    sc = SCENES[name]()
    tree = scenes.build_tree(sc, S.Octree)
    t0 = time.perf_counter()
    host = S.OctreeGPUHost(tree)
    full_ms = (time.perf_counter() - t0) * 1e3
    out = {"scene": sc.name, "full_upload_ms": full_ms, "full": host.last_upload(), "stats": host.stats(), "edits": []}
    rng = np.random.default_rng(3)
    n = sc.tree_size
    for k in (1, 16, 256, 4096):
        pos = rng.integers(0, n, (k, 3))
        pos[:, 1] = rng.integers(0, n // 4, k)  # stay near the ground where bricks exist
        t0 = time.perf_counter()
        for p in pos:
            tree.insert(tuple(int(v) for v in p), (200, 50, 50, 255))
        edit_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        host.reload()
        reload_ms = (time.perf_counter() - t0) * 1e3
        out["edits"].append({"inserts": k, "insert_ms": edit_ms, "reload_ms": reload_ms, **host.last_upload()})
        print(f"{sc.name}: {k:5d} inserts {edit_ms:8.2f} ms, reload {reload_ms:8.2f} ms, {host.last_upload()}", flush=True)
    full = out["full"]
    algo = full["bricks"] * (sc.brick_dim ** 3) * (4 + 1 / 8)  # 4 B read per voxel, 1 bit written
    gbs = algo / (full["bits_kernel_ms"] * 1e-3) / 1e9 if full["bits_kernel_ms"] > 0 else 0.0
    out["bits_kernel_gbs"] = gbs
    print(f"{sc.name}: full upload {full_ms:.1f} ms {full}; occupancy_bits_kernel {full['bits_kernel_ms']:.4f} ms = {gbs:.0f} GB/s algorithmic")
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"reload_probe_{name}.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
