#!/usr/bin/env python3
"""Randomised GPU-vs-oracle parity soak (GPU box): random trees (sizes, brick dims, edit mixes incl. insert_at_lod and
clear), random rays (outside / inside / axis-parallel / grazing), every output field compared bit for bit.

    python tools/fuzz_parity.py [--seconds 120] [--seed 0] [--mips] [--host-mirror]
--mips: every tree gets MIP maps (random strategy, switched on before, midway or after the edits) and the rays are
queried through get_by_ray_at_lod at random viewing distances (src/raytracing/raytracing_on_cpu.rs:325).
--host-mirror: no GPU - the rays go through the host build of the kernels' source (tests/host_mirror: traverse.cuh compiled by
g++) instead of the device; checks the logic of the GPU path, not what nvcc / ptxas make of it.
Writes gpurun_out/fuzz_parity.json. Exit code 1 on the first mismatch (the failing case is printed).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    sys.path.insert(0, p)
import oracle_lib as O  # noqa: E402
import shocovox_b200 as S  # noqa: E402
from product_adapter import ProductOctree  # noqa: E402


def random_tree_ops(rng, size, dim):
    ops = []
    n = int(rng.integers(20, 1500))
    colors = [int(c) for c in rng.integers(1, 2 ** 32 - 1, int(rng.integers(1, 6)), dtype=np.uint64) | 0xFF]
    style = int(rng.integers(0, 4))
    for _ in range(n):
        if style == 0:      # scattered voxels
            p = rng.integers(0, size, 3)
        elif style == 1:    # a slab / floor
            p = np.array([rng.integers(0, size), rng.integers(0, max(size // 8, 1)), rng.integers(0, size)])
        elif style == 2:    # clustered blob
            p = np.clip(rng.normal(size / 2, size / 8, 3), 0, size - 1).astype(np.int64)
        else:               # mix with lod inserts and clears
            p = rng.integers(0, size, 3)
        p = tuple(int(v) for v in p)
        k = int(rng.integers(0, 12))
        c = colors[int(rng.integers(0, len(colors)))]
        if style == 3 and k < 2:
            lod = int(2 ** rng.integers(1, max(2, int(np.log2(size)) - 1)))
            ops.append(("insert_at_lod", tuple((v // lod) * lod for v in p), lod, c))
        elif style == 3 and k < 4:
            ops.append(("clear", p))
        elif k == 11:
            ops.append(("insert_data", p, int(rng.integers(1, 5))))
        else:
            ops.append(("insert", p, c))
    return ops


def apply(t, ops):
    for op in ops:
        if op[0] == "insert":
            t.insert(op[1], op[2])
        elif op[0] == "insert_data":
            t.insert(op[1], None, op[2])
        elif op[0] == "insert_at_lod":
            t.insert_at_lod(op[1], op[2], op[3])
        else:
            t.clear(op[1])


def random_rays(rng, size, n):
    origin = rng.uniform(-1.5 * size, 2.5 * size, (n, 3)).astype(np.float32)
    inside = rng.random(n) < 0.3
    origin[inside] = rng.uniform(0, size, (int(inside.sum()), 3)).astype(np.float32)
    target = rng.uniform(-0.1 * size, 1.1 * size, (n, 3)).astype(np.float32)
    d = target - origin
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    d = (d / ln[:, None]).astype(np.float32)
    k = n // 10
    d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], k)[:, None].astype(np.float32)
    origin[:k] = np.round(origin[:k])
    j = n // 10  # grazing: along a face / grid plane
    origin[k:k + j, 1] = rng.integers(0, size + 1, j).astype(np.float32)
    d[k:k + j, 1] = 0.0
    ln = np.sqrt((d[k:k + j, 0] ** 2 + d[k:k + j, 1] ** 2) + d[k:k + j, 2] ** 2, dtype=np.float32)
    d[k:k + j] = (d[k:k + j] / np.maximum(ln, 1e-12)[:, None]).astype(np.float32)
    return np.concatenate([origin, d], axis=1)


def bits(a):
    """f32 bit patterns, with every NaN mapped to one canonical pattern (0/0 gives 0xFFC00000 on x86 and 0x7FFFFFFF
    on the GPU; a NaN's sign and payload carry no meaning - the reference's normal of a hit at the exact cell centre)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.where(np.isnan(a), np.uint32(0x7FC00000), a.view(np.uint32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--rays", type=int, default=20000)
    ap.add_argument("--mips", action="store_true")
    ap.add_argument("--host-mirror", action="store_true")
    ap.add_argument("--mirror-define", action="append", default=[], help="with --host-mirror: build the mirror with -D<this> (e.g. SVX_SHARED_RCP=1)")
    args = ap.parse_args()
    mirror = None
    if args.host_mirror:
        import test_host_mirror as HM

        tag = "".join(c if c.isalnum() else "_" for c in "_".join(args.mirror_define))
        mirror = HM.build_mirror("host_mirror" + ("_" + tag if tag else ""), args.mirror_define)
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    cases = rays_total = hits_total = mip_probes_total = 0
    configs = [(4, 1), (8, 1), (8, 2), (16, 2), (16, 4), (32, 4), (32, 8), (64, 8), (64, 16), (128, 8), (128, 32), (256, 8), (512, 8), (1024, 16), (32, 1)]
    while time.time() - t0 < args.seconds:
        size, dim = configs[int(rng.integers(0, len(configs)))]
        if args.mips and size > 256:
            continue
        ops = random_tree_ops(rng, size, dim)
        a, b = O.OracleOctree(size, dim), ProductOctree(size, dim)
        vds = [S.F32_MAX]
        if args.mips:
            for lvl in range(1, 6):
                m = int(rng.integers(0, 6))
                if m == 2 and lvl > 3:
                    m = 1  # PointFilterBD samples (2^level)^3 voxels per MIP voxel: keep the soak moving
                if m < 5:
                    thr = float(rng.choice([0.0, 0.05, 0.2, 0.5, 1.0]))
                    for t in (a, b):
                        t.set_method_at(lvl, m, thr)
                if rng.random() < 0.3:
                    thr = float(rng.choice([0.0, 0.01, 0.1, 0.4]))
                    for t in (a, b):
                        t.set_color_similarity_thr_at(lvl, thr)
            when = int(rng.integers(0, 3))  # MIPs on: before the edits / midway / after (one recalculation)
            cut = 0 if when == 0 else (len(ops) // 2 if when == 1 else len(ops))
            for t in (a, b):
                apply(t, ops[:cut])
                t.switch_albedo_mip_maps(True)
                apply(t, ops[cut:])
            if a.mip_hash() != b.mip_hash():
                print("MIP MISMATCH", size, dim, when, ops[:20])
                return 1
            vds = [S.F32_MAX] + [float(v) for v in rng.choice([0.5, 3.0, 10.0, 40.0, 150.0, 600.0, 4000.0], 2, replace=False)]
        else:
            apply(a, ops)
            apply(b, ops)
        if a.structure_hash() != b.structure_hash():
            print("TREE SHAPE MISMATCH", size, dim, ops[:20])
            return 1
        host = None if mirror else S.OctreeGPUHost(b.tree)
        for vd in vds:
            rays = random_rays(rng, size, args.rays)
            o = a.get_by_rays_at_lod(rays, vd)
            if mirror:  # the host build of the kernel code; it returns the palette value, not the resolved entry
                m = HM.mirror_rays(mirror, b.tree, rays, specialise=int(rng.integers(0, 2)), viewing_distance=vd if args.mips else None)
                g = np.zeros(len(rays), dtype=o.dtype)
                for k in ("rgba", "data", "entry_kind"):
                    g[k] = o[k]
                g["hit"], g["palette_value"] = m["hit"], m["palette_value"]
                g["impact_point"], g["normal"], g["distance"] = m["impact_point"], m["normal"], m["distance"]
            else:
                g = host.get_by_rays(rays, vd)
            ok = (np.array_equal(g["hit"] != 0, o["hit"] != 0) and np.array_equal(g["palette_value"], o["palette_value"])
                  and np.array_equal(g["rgba"], o["rgba"]) and np.array_equal(g["data"], o["data"])
                  and np.array_equal(bits(g["impact_point"]), bits(o["impact_point"]))
                  and np.array_equal(bits(g["normal"]), bits(o["normal"])) and np.array_equal(bits(g["distance"]), bits(o["distance"])))
            if not ok:
                fields = {
                    "hit": g["hit"] != o["hit"], "palette_value": g["palette_value"] != o["palette_value"],
                    "entry_kind": g["entry_kind"] != o["entry_kind"],
                    "rgba": (g["rgba"] != o["rgba"]).any(axis=1), "data": g["data"] != o["data"],
                    "impact_point": (bits(g["impact_point"]) != bits(o["impact_point"])).any(axis=1),
                    "normal": (bits(g["normal"]) != bits(o["normal"])).any(axis=1),
                    "distance": bits(g["distance"]) != bits(o["distance"]),
                }
                print("RAY MISMATCH size", size, "dim", dim, "case", cases, {k: int(v.sum()) for k, v in fields.items()})
                for k, v in fields.items():
                    idx = np.nonzero(v)[0][:3]
                    for i in idx:
                        print(" ", k, "ray", rays[i].tolist(), "gpu", g[i], "oracle", {n: o[i][n] for n in g.dtype.names})
                return 1
            if int(o["would_panic"].sum()) and not args.mips:  # with MIPs a miss leaves the point outside the node: common
                print("note: reference would have panicked on", int(o["would_panic"].sum()), "rays in case", cases)
            rays_total += len(rays)
            hits_total += int(o["hit"].sum())
            mip_probes_total += int(o["mip_probes"].sum())
        cases += 1
    out = {"seed": args.seed, "seconds": round(time.time() - t0, 1), "random_trees": cases, "rays": rays_total, "hits": hits_total,
           "mips": bool(args.mips), "mip_probes": mip_probes_total, "through": ("host mirror (CPU)" + (" built with " + " ".join(args.mirror_define) if args.mirror_define else "")) if mirror else "GPU",
           "result": "every field bit-identical to the CPU oracle"}
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"fuzz_parity_{'host_mirror_' if mirror else ''}{'mips_' if args.mips else ''}seed{args.seed}.json").write_text(json.dumps(out))
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
