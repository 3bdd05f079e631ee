#!/usr/bin/env python3
"""Freeze the CPU oracle's outputs on the reference's deterministic cases into tests/golden/oracle_outputs.json.

The reference (Rust) cannot be run here, so these are NOT reference outputs; they pin the oracle itself: the 17 literal
edge-case rays of src/raytracing/tests.rs:253-813 (hit flag, palette value, impact point / normal / distance bit
patterns, loop counters) and SHA-256 digests of whole oracle frames of the examples/cpu_render.rs scene. A later change
that alters the oracle and the kernel together can no longer go unnoticed (tests/test_oracle_golden.py).
The survey's independent emulation of the reference recorded impact x = 511.00104 and 2 outer iterations for
`deep_stack` and 120 restarts for `cube_flaps`; both are reproduced in this file.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    sys.path.insert(0, p)
import oracle_lib as O  # noqa: E402
from ray_cases import CASES  # noqa: E402
from shocovox_b200 import scenes  # noqa: E402


def bits(v):
    return [int(x) for x in np.asarray(v, dtype=np.float32).view(np.uint32).ravel()]


def main():
    out = {"rays": {}, "frames": {}}
    for c in CASES:
        t = O.OracleOctree(c["size"], c["dim"])
        c["build"](t)
        h = t.get_by_ray(c["origin"], c["direction"])
        out["rays"][c["name"]] = {
            "reference_line": c["line"], "hit": int(h.hit), "palette_value": int(h.palette_value),
            "impact_point_bits": bits(h.impact_point[:]), "normal_bits": bits(h.normal[:]), "distance_bits": bits([h.distance])[0],
            "node_iters": int(h.node_iters), "voxel_fetches": int(h.voxel_fetches), "outer_iters": int(h.outer_iters),
            "crawl_iters": int(h.crawl_iters), "structure_hash": int(t.structure_hash()),
        }
    scene = scenes.cpu_render_scene()
    tree = scenes.build_tree(scene, O.OracleOctree)
    out["frames"]["scene"] = {"name": scene.name, "voxels": int(len(scene.xyz)), "structure_hash": int(tree.structure_hash())}
    for k in (0, 21, 63):
        cam = scenes.cpu_render_camera(64, k)
        f = tree.render(O.make_camera(cam.origin, cam.direction, cam.frustum[0], cam.frustum[1], cam.glass_distance), 150, 150)
        out["frames"][f"cpu_render_150_k{k}"] = {
            "camera_origin_bits": bits(cam.origin), "camera_direction_bits": bits(cam.direction),
            "hits": int((f["hit_id"] != 0xFFFFFFFF).sum()),
            "hit_id_sha256": hashlib.sha256(f["hit_id"].tobytes()).hexdigest(),
            "albedo_sha256": hashlib.sha256(f["albedo"].tobytes()).hexdigest(),
            "distance_sha256": hashlib.sha256(f["distance"].tobytes()).hexdigest(),
            "node_iters": f["node_iters"], "voxel_fetches": f["voxel_fetches"], "outer_iters": f["outer_iters"],
        }
    path = ROOT / "tests" / "golden" / "oracle_outputs.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
