#!/usr/bin/env python3
"""Primary-ray throughput of the B200 path on the reference's named configs (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                    [--mode auto|tiles_fused|tiles_nccl|poses] [--wire 8|12] [--no-extra]

A step = one pass of the hot path over one batch: ONE full frame (in-kernel ray generation + get_by_ray per pixel +
framebuffer write). Default workload = BASELINE configs[3], the sponza-scale mixed-resolution tree at 3840x2160 (the
configuration the north-star target is quoted on: "at 4K", "screen-tile sharded over 2/4/8").
N = 1: one GPU renders the frame. N > 1 (torchrun, one rank per GPU): the SAME frame is split into interleaved bands of 8
image rows, every rank renders its bands from its own replica of the tree, and the traversal kernels of ranks 1..N-1 store
their pixels straight into rank 0's framebuffer over NVLink (CUDA IPC; device-side go / done flags, no host barrier, no
collective per frame: csrc/multi_gpu.cu). The step time is rank 0's: its own kernel plus the wait until the slowest peer's
rows have arrived -> strong scaling. `--mode tiles_nccl` gathers compact bands with NCCL instead (comparison), `--mode poses`
shards camera poses (BASELINE configs[4]; no exchange, weak scaling).
extra_workloads (on by default): N = 1: configs[1] dot_cube 1080p, configs[2] minecraft-style 4K, configs[4] terrain pose
1080p, each with ms/frame, roofline fraction and parity against the oracle; N > 1: configs[4], all 256 poses pose-sharded.

Prints ONE JSON line (rank 0). `value` is device-timed with CUDA events on the launching stream, L2 flushed before
every timed step; `e2e` goes through the public API with host buffers (pose in, framebuffer out, copies timed).
Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "primary_mrays_per_s"
UNIT = "Mrays/s"

# name: (scene kind, camera kwargs, (w, h), description)
WORKLOADS = {
    "dot_cube_1080p": ("dot_cube", dict(zoom=True), (1920, 1080),
                       "BASELINE configs[1]: examples/dot_cube.rs tree 256/32, camera (512,128,-512)->0, glass 10x10 at "
                       "frustum.z=200 (the example's own CPU ray loop, dot_cube.rs:204-231), 1920x1080"),
    "dot_cube_1080p_fov": ("dot_cube", dict(zoom=False), (1920, 1080),
                           "examples/dot_cube.rs tree 256/32, glass 10x10 at fov=3 (the shader's placement), 1920x1080"),
    "dot_cube_4k": ("dot_cube", dict(zoom=True), (3840, 2160), "examples/dot_cube.rs tree 256/32, glass at frustum.z, 3840x2160"),
    "cpu_render_150": ("cpu_render", {}, (150, 150), "BASELINE configs[0]: examples/cpu_render.rs tree 64/8, 150x150"),
    "cpu_render_1080p": ("cpu_render", {}, (1920, 1080), "examples/cpu_render.rs tree 64/8, 1920x1080"),
    "cpu_render_4k": ("cpu_render", {}, (3840, 2160), "examples/cpu_render.rs tree 64/8, 3840x2160"),
    "minecraft_4k": ("minecraft", {}, (3840, 2160),
                     "BASELINE configs[2]: synthetic minecraft-style blocky heightfield 1024/32 (minecraft.vox is not in the checkout), 3840x2160"),
    "sponza_4k": ("sponza", {}, (3840, 2160),
                  "BASELINE configs[3]: synthetic sponza-scale colonnade 2048/32, insert_at_lod slabs + per-voxel detail (mixed-resolution bricks), 3840x2160"),
    "terrain_poses_1080p": ("terrain_poses", {}, (1920, 1080),
                            "BASELINE configs[4]: 256 camera poses orbiting a 1024/8 value-noise terrain, 1920x1080 per pose, pose k -> rank k mod N"),
}
HEAVY = {"minecraft_4k", "sponza_4k", "terrain_poses_1080p"}
BAND_ROWS = 8  # rows per band of the tile-sharded frame  # the CPU leg samples rows instead of whole frames


def make_workload(name: str):
    """-> (scene, [camera per pose], (w, h), description)"""
    from shocovox_b200 import scenes

    kind, cam_kw, res, desc = WORKLOADS[name]
    if kind == "dot_cube":
        return scenes.dot_cube_scene(), [scenes.dot_cube_camera(**cam_kw)], res, desc
    if kind == "cpu_render":
        return scenes.cpu_render_scene(), [scenes.cpu_render_camera()], res, desc
    if kind == "minecraft":
        return scenes.terrain_scene(1024, 32, 1234, 4, shell=8, name="minecraft"), [scenes.terrain_camera(1024)], res, desc
    if kind == "sponza":
        return scenes.colonnade_scene(2048, 32), [scenes.colonnade_camera(2048)], res, desc
    if kind == "terrain_poses":
        return scenes.terrain_scene(1024, 8, 4321, 1, shell=4), scenes.orbit_cameras(1024, 256), res, desc
    raise KeyError(name)


# ---- clocks -------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU from a thread while the timed regions run (NVML)."""

    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop, self._active = threading.Event(), threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.err = str(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.reasons.update(name for bit, name in self.REASONS.items() if mask & bit)
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t.start()

    def window(self, on: bool):
        (self._active.set if on else self._active.clear)()

    def stop(self) -> dict:
        self._stop.set()
        if self.ok and self.t.is_alive():
            self.t.join(timeout=1)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": getattr(self, "err", "no samples inside the timed region")}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---- CPU oracle legs -------------------------------------------------------------------------------------------------
F32_MAX = 3.4028234663852886e38


def viewing_distance_of(args, cam) -> float:
    """--mips VD: get_by_ray_at_lod's viewing distance; "frustum" = the camera's viewport.frustum.z, which is what the
    reference's own GPU path feeds (assets/shaders/viewport_render.wgsl:631-638). Without --mips: f32::MAX (get_by_ray)."""
    if args.mips is None:
        return F32_MAX
    return float(cam.frustum[2]) if args.mips == "frustum" else float(args.mips)


def enable_mips(args, product_tree=None, oracle_tree=None):
    """--mips: MIPMapStrategy::default() switched on after construction (one recalculate_mips), on both implementations."""
    if args.mips is None:
        return 0.0
    t0 = time.time()
    if product_tree is not None:
        product_tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
    if oracle_tree is not None:
        oracle_tree.switch_albedo_mip_maps(True)
    return time.time() - t0


def oracle_camera(cam):
    import oracle_lib as O

    return O.make_camera(cam.origin, cam.direction, cam.frustum[0], cam.frustum[1], cam.glass_distance)


def sample_rows(height: int, n_rows: int) -> np.ndarray:
    """n_rows image rows spread evenly over the frame (sky and ground are both represented)."""
    n_rows = max(1, min(n_rows, height))
    return np.unique(np.linspace(0, height - 1, n_rows).round().astype(np.uint32))


def oracle_timed_sample(otree, cam, res, budget_s: float, whole_frames: bool, vd: float = F32_MAX):
    """Times the CPU oracle with all host threads on the workload: whole frames when they are quick, else a bounded
    sample of evenly spread rows sized from a probe so that the leg takes about `budget_s` seconds."""
    import oracle_lib as O

    ocam = oracle_camera(cam)
    threads = int(O.lib().svxo_hardware_threads())
    w, h = res
    if whole_frames:
        frames, total = [], 0.0
        while len(frames) < 12 and (total < budget_s or len(frames) < 2):
            f = otree.render(ocam, w, h, threads=threads, viewing_distance=vd)
            frames.append(f)
            total += f["seconds"]
        best = min(frames, key=lambda f: f["seconds"])
        rows = np.arange(h, dtype=np.uint32)
        return {"frame": best, "rows": rows, "threads": threads, "rays": w * h, "seconds": best["seconds"],
                "mrays": w * h / best["seconds"] / 1e6,
                "sample": f"{len(frames)} whole frames of the workload ({w * h} rays each) on {threads} host threads; best frame"}
    probe_rows = sample_rows(h, 4)
    probe = otree.render(ocam, w, h, threads=threads, row_list=probe_rows, viewing_distance=vd)
    per_row = max(probe["seconds"] / len(probe_rows), 1e-6)
    n = int(min(h, max(8, budget_s / per_row)))
    rows = sample_rows(h, n)
    f = otree.render(ocam, w, h, threads=threads, row_list=rows, viewing_distance=vd)
    rays = len(rows) * w
    return {"frame": f, "rows": rows, "threads": threads, "rays": rays, "seconds": f["seconds"], "mrays": rays / f["seconds"] / 1e6,
            "sample": f"{len(rows)} of {h} image rows spread evenly over one frame ({rays} rays) on {threads} host threads"}


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU get_by_ray loop. The Rust crate cannot be built here, so this is the
    oracle port, with all host threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return 0
    import oracle_lib as O
    from shocovox_b200 import scenes

    scene, cams, res, desc = make_workload(args.workload)
    tree = scenes.build_tree(scene, O.OracleOctree)
    enable_mips(args, oracle_tree=tree)
    vd = viewing_distance_of(args, cams[0])
    threads = int(O.lib().svxo_hardware_threads())
    w, h = res
    whole = args.workload not in HEAVY
    rows = np.arange(h, dtype=np.uint32)
    if not whole:
        probe = tree.render(oracle_camera(cams[0]), w, h, threads=threads, row_list=sample_rows(h, 4), viewing_distance=vd)
        per_row = max(probe["seconds"] / 4, 1e-6)
        budget = 120.0 / max(args.steps + args.warmup, 1)
        rows = sample_rows(h, int(min(h, max(4, budget / per_row))))
    for i in range(args.warmup):
        tree.render(oracle_camera(cams[i % len(cams)]), w, h, threads=threads, row_list=rows, viewing_distance=vd)
    t = 0.0
    for i in range(args.steps):
        t += tree.render(oracle_camera(cams[i % len(cams)]), w, h, threads=threads, row_list=rows, viewing_distance=vd)["seconds"]
    rays = len(rows) * w
    value = rays * args.steps / t / 1e6
    sample = (f"{args.steps} steps, each {'one whole frame' if whole else f'{len(rows)} of {h} rows spread over the frame'} "
              f"({rays} rays) on {threads} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.mode == "poses" else "strong",  # as our arm labels the same launch
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the workload is our arm's: one whole frame per step; this arm times a bounded sample of its rows (cpu_baseline.sample)
        "config": {"workload": args.workload, "description": desc, "resolution": list(res), "rays_per_step": w * h,
                   "rays_sampled_per_step": rays,
                   **({"mips": {"strategy": "MIPMapStrategy::default(), enabled", "viewing_distance": vd}} if args.mips is not None else {})},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU get_by_ray of shocovox-rs restated in C++ (oracle/): the Rust crate cannot be compiled in this image (no cargo/rustc)",
    }))
    return 0


# ---- our arm ---------------------------------------------------------------------------------------------------------
def peak_hbm():
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        return float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def kernel_name(scene, mips: bool) -> str:
    # kernels.cu: launch_render picks the instantiation for the tree's brick dimension (8 / 32: compile-time strides)
    return ("svx::render_lod_kernel" if mips else "svx::render_kernel") + {8: "_brick8", 32: "_brick32"}.get(int(scene.brick_dim), "")


def oracle_leg(name, scene, cam, res, vd, args, ms_per_launch, rays_per_launch, tree_bytes, check_view, mips_tree=None, budget=None):
    """CPU oracle on a bounded sample of the workload: cpu_baseline, the algorithmic bytes of one launch (roofline) and a
    parity check of the frame `check_view` renders (outside every timed region). ms_per_launch / rays_per_launch describe the
    launch of the dominant kernel that was timed (a whole frame, or this rank's share of it)."""
    import oracle_lib as O
    from shocovox_b200 import scenes

    peak, peak_src = peak_hbm()
    w, h = res
    rays_per_frame = w * h
    t0 = time.time()
    otree = scenes.build_tree(scene, O.OracleOctree)
    enable_mips(args, oracle_tree=otree)
    t_obuild = time.time() - t0
    o = oracle_timed_sample(otree, cam, res, budget if budget is not None else args.cpu_budget, whole_frames=name not in HEAVY, vd=vd)
    f = o["frame"]
    scale = rays_per_frame / o["rays"]
    alg_bytes = 12 * rays_per_frame + (16 * f["node_iters"] + 4 * f["voxel_fetches"]) * scale  # one whole frame
    share = rays_per_launch / rays_per_frame
    t_kernel = ms_per_launch * 1e-3
    achieved = alg_bytes * share / t_kernel / 1e9
    alg_bytes_nocrawl = alg_bytes - 16 * f["crawl_iters"] * scale
    achieved_nocrawl = alg_bytes_nocrawl * share / t_kernel / 1e9
    traffic = None
    tpath = ROOT / "profiles" / "ncu_traffic.json"
    if tpath.exists():
        traffic = json.loads(tpath.read_text()).get(name)
    out = {}
    out["roofline"] = {
        # the primary figures charge only what the kernel cannot skip: restarts it resolves in closed form are left out
        "bound": "hbm", "achieved": achieved_nocrawl, "peak": peak, "unit": "GB/s", "frac": achieved_nocrawl / peak, "traffic": traffic,
        "achieved_task_definition": achieved, "frac_task_definition": achieved / peak,
        "achieved_excluding_fast_forwarded_restarts": achieved_nocrawl, "frac_excluding_fast_forwarded_restarts": achieved_nocrawl / peak,
        "peak_source": peak_src, "kernel": kernel_name(scene, args.mips is not None),
        "kernel_ms_per_launch": ms_per_launch, "rays_per_launch": rays_per_launch,
        "algorithmic_bytes_per_launch": alg_bytes_nocrawl * share, "algorithmic_bytes_per_launch_task_definition": alg_bytes * share,
        "per_ray": {"node_visits": f["node_iters"] / o["rays"], "voxel_fetches": f["voxel_fetches"] / o["rays"],
                    "restarts": f["outer_iters"] / o["rays"], "crawl_restarts": f["crawl_iters"] / o["rays"],
                    "bytes": alg_bytes / rays_per_frame, "rays_entering_root": f["rays_in_root"] / o["rays"]},
        "compulsory_bound_ms": (tree_bytes + 12 * rays_per_frame) / (peak * 1e9) * 1e3,
        "note": "B_ray = 12 + 16 N_node + 4 N_vox, N counted by the CPU oracle executing the reference algorithm on "
                + ("the same rays" if scale == 1 else "a row sample of the same frame, scaled")
                + " (SURVEY 8(d)). `frac` / `achieved` leave out the 16 B of every 0.1-nudge restart in which the reference pops the root"
                  " right away (crawl_restarts per ray): the kernel applies those in closed form, bit-exactly, without touching memory, so"
                  " charging them would credit bytes that never move. *_task_definition charges them as SURVEY 8(d) literally says and exceeds"
                  " 1 on crawl-heavy scenes for that reason. ncu: the kernel is instruction-issue bound over an L1/L2-resident tree, DRAM traffic is small",
    }
    out["cpu_baseline"] = {"value": o["mrays"], "unit": UNIT, "cores": o["threads"], "kind": "port", "sample": o["sample"],
                           "oracle_tree_build_s": round(t_obuild, 2)}
    # the reference's own caller loop is single-threaded (examples/cpu_render.rs:104-105): the same on ONE host thread
    try:  # an auxiliary figure: it must never cost the bench line
        per_row_1t = o["seconds"] / len(o["rows"]) * o["threads"]
        rows_1t = sample_rows(h, int(min(h, max(8, 2.0 / max(per_row_1t, 1e-6)))))
        f1 = otree.render(oracle_camera(cam), w, h, threads=1, row_list=rows_1t, viewing_distance=vd)
        out["cpu_baseline"]["single_thread_value"] = len(rows_1t) * w / f1["seconds"] / 1e6
        out["cpu_baseline"]["single_thread_sample"] = f"{len(rows_1t)} of {h} image rows ({len(rows_1t) * w} rays) on 1 host thread"
    except Exception as e:  # noqa: BLE001
        out["cpu_baseline"]["single_thread_value"] = None
        out["cpu_baseline"]["single_thread_sample"] = f"failed: {e}"
    got = check_view() if callable(check_view) else check_view
    rows = o["rows"]
    out["parity_vs_oracle"] = {
        "rows_checked": int(len(rows)),
        "hit_id_equal": bool(np.array_equal(got["hit_id"][rows], f["hit_id"][rows])),
        "albedo_equal": bool(np.array_equal(got["albedo"][rows], f["albedo"].view(np.uint32)[..., 0][rows])),
        "distance_bits_equal": bool(np.array_equal(got["distance"][rows].view(np.uint32), f["distance"][rows].view(np.uint32))),
        "would_panic": int(f["would_panic"]),
        **({"mip_probes": int(f["mip_probes"]), "mip_hash_equal": bool(mips_tree.albedo_mip_map_resampling_strategy().mip_hash() == otree.mip_hash())}
           if (args.mips is not None and mips_tree is not None) else {}),
    }
    return out


def timed_frames(view, vps, n, warm=3):
    """ms per frame of `n` L2-flushed renders cycling through the poses `vps` (after `warm` untimed ones)."""
    ms = []
    for i in range(warm + n):
        view.set_viewport(vps[i % len(vps)])
        view.flush_l2()
        k = view.render(sync=True)["kernel_ms"]
        if i >= warm:
            ms.append(k)
    return ms


def extra_single_gpu(name, args, local_rank):
    """One more BASELINE config on this GPU: kernel ms/frame (L2 flushed), roofline fraction and parity against the oracle."""
    import shocovox_b200 as S
    from shocovox_b200 import scenes

    scene, cams, res, desc = make_workload(name)
    tree = scenes.build_tree(scene, S.Octree)
    host = S.OctreeGPUHost(tree, local_rank)
    vps = [S.Viewport(c.origin, c.direction, c.frustum, c.fov) for c in cams]
    view = host.create_new_view(64, vps[0], res)
    if cams[0].glass_at_frustum_z:
        view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
    ms0 = timed_frames(view, vps[:1], 10)
    out = {"description": desc, "resolution": list(res), "ms_per_frame": float(np.mean(ms0)),
           "mrays_per_s": res[0] * res[1] / (np.mean(ms0) * 1e-3) / 1e6, "tree_bytes": host.stats()["total_bytes"]}
    if len(vps) > 1:  # a pose batch: 8 poses spread over the orbit
        spread = vps[:: max(1, len(vps) // 8)][:8]
        msp = timed_frames(view, spread, 16)
        out["poses_sampled"] = len(spread)
        out["ms_per_frame_over_poses"] = float(np.mean(msp))
        out["mrays_per_s_over_poses"] = res[0] * res[1] / (np.mean(msp) * 1e-3) / 1e6
        view.set_viewport(vps[0])
    if not args.no_cpu_baseline:
        leg = oracle_leg(name, scene, cams[0], res, F32_MAX, args, float(np.mean(ms0)), res[0] * res[1], out["tree_bytes"],
                         view.render_to_host, budget=min(args.cpu_budget, 4.0))
        out["roofline_frac"] = leg["roofline"]["frac"]
        out["roofline"] = {k: leg["roofline"][k] for k in ("achieved", "peak", "frac", "frac_task_definition", "kernel",
                                                            "algorithmic_bytes_per_launch", "per_ray", "traffic")}
        out["cpu_baseline"] = leg["cpu_baseline"]
        out["parity_vs_oracle"] = leg["parity_vs_oracle"]
    return out


def extra_pose_batch(args, rank, local_rank, world, dist, torch):
    """BASELINE configs[4] on all ranks: 256 poses over the 1024/8 terrain, pose k -> rank k mod N, no exchange."""
    import shocovox_b200 as S
    from shocovox_b200 import scenes

    name = "terrain_poses_1080p"
    scene, cams, res, desc = make_workload(name)
    tree = scenes.build_tree(scene, S.Octree)
    host = S.OctreeGPUHost(tree, local_rank)
    vps = [S.Viewport(c.origin, c.direction, c.frustum, c.fov) for c in cams]
    view = host.create_new_view(64, vps[0], res)
    mine = list(range(rank, len(vps), world))
    for k in mine[:3]:
        view.set_viewport(vps[k])
        view.render(sync=True)
    view.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    total = 0.0
    for k in mine:
        view.set_viewport(vps[k])
        view.flush_l2()
        total += view.render(sync=True)["kernel_ms"]
    t = torch.tensor([total], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total = float(t.item())
    out = {"description": desc, "resolution": list(res), "poses": len(vps), "parallelism": "pose-sharded x%d (tree replicated, no exchange)" % world,
           "ms_per_batch": total, "ms_per_pose_per_gpu": total / max(len(mine), 1), "mrays_per_s": len(vps) * res[0] * res[1] / (total * 1e-3) / 1e6}
    if rank == 0 and not args.no_cpu_baseline:
        import oracle_lib as O

        otree = scenes.build_tree(scene, O.OracleOctree)
        rows = sample_rows(res[1], 32)
        ok = {"hit_id_equal": True, "albedo_equal": True, "distance_bits_equal": True}
        checked = [k for k in (0, 64, 128, 192) if k < len(vps)]
        for k in checked:
            view.set_viewport(vps[k])
            got = view.render_to_host()
            f = otree.render(oracle_camera(cams[k]), res[0], res[1], threads=int(O.lib().svxo_hardware_threads()), row_list=rows)
            ok["hit_id_equal"] &= bool(np.array_equal(got["hit_id"][rows], f["hit_id"][rows]))
            ok["albedo_equal"] &= bool(np.array_equal(got["albedo"][rows], f["albedo"].view(np.uint32)[..., 0][rows]))
            ok["distance_bits_equal"] &= bool(np.array_equal(got["distance"][rows].view(np.uint32), f["distance"][rows].view(np.uint32)))
        out["parity_vs_oracle"] = {"poses_checked": checked, "rows_per_pose": int(len(rows)), **ok}
    return name, out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sponza_4k", choices=list(WORKLOADS))
    ap.add_argument("--mode", default="auto", choices=["auto", "poses", "tiles_nccl", "tiles_fused"],
                    help="N > 1: auto = tiles_fused (one frame, row bands, fused gather into rank 0's framebuffer)")
    ap.add_argument("--wire", default="auto", choices=["auto", "8", "12"],
                    help="tiles_fused: bytes per pixel the peers send (12 = hit id, albedo, distance; 8 = hit id and distance, rank 0 "
                         "resolves albedo = palette[hit id] for the received rows). auto = 12: beyond 4 GPUs the peers write whole 128-byte rows "
                         "from a shared-memory stage, which NVLink delivers fast enough (profiles/r02_gather_probe_n8_staged.json)")
    ap.add_argument("--mips", default=None, metavar="VD",
                    help="switch the tree's MIP maps on and render through get_by_ray_at_lod at this viewing distance "
                         "(a number, or 'frustum' for the camera's viewport.frustum.z like the reference's shader)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=8.0, help="seconds of CPU-oracle work for the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_workloads (the other BASELINE configs)")
    ap.add_argument("--extra", action="store_true", help=argparse.SUPPRESS)  # round-1 flag, accepted and ignored
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import shocovox_b200 as S
    from shocovox_b200 import distributed as D, scenes

    if S.cuda_device_count() < 1:
        print(json.dumps({"error": "no CUDA device: the ray path has no CPU fallback"}))
        return 1
    dist = torch = None
    if world > 1:
        # keep this rank's host threads and first-touch pinned pages on the CPU cores NVML reports as local to its GPU
        try:
            import pynvml

            pynvml.nvmlInit()
            words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank), (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
            if cpus:
                os.sched_setaffinity(0, cpus & set(range(os.cpu_count())) or cpus)
        except Exception:
            pass
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mode = args.mode
    if world == 1:
        mode = "single"
    elif mode == "auto":
        mode = "tiles_fused"
    tiles = mode in ("tiles_fused", "tiles_nccl")
    wire_bytes = 12 if args.wire == "auto" else int(args.wire)
    wire = S.WIRE_ID_DISTANCE if wire_bytes == 8 else S.WIRE_THREE_PLANES

    scene, cams, res, desc = make_workload(args.workload)
    w, h = res
    t_build = time.time()
    tree = scenes.build_tree(scene, S.Octree)
    t_mips = enable_mips(args, product_tree=tree)
    vd = viewing_distance_of(args, cams[0])
    host = S.OctreeGPUHost(tree, local_rank)
    t_build = time.time() - t_build
    vps = [S.Viewport(c.origin, c.direction, c.frustum, c.fov) for c in cams]

    def new_view():
        v = host.create_new_view(64, vps[0], res)
        if cams[0].glass_at_frustum_z:
            v.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
        v.set_viewing_distance(vd)
        return v

    view = new_view()
    rays_per_frame = w * h
    # pose mode: step i of rank r renders pose (r + i * world) mod n_poses ; otherwise everybody renders pose i
    def pose_index(i):
        return (rank + i * world if mode == "poses" else i) % len(vps)

    gather = None
    if mode == "tiles_fused":
        D.open_gather(view, rank, world, BAND_ROWS, wire)
    elif mode == "tiles_nccl":
        view.set_shard(rank, world, BAND_ROWS)
        view.set_compact_rows(True)
        lr = D.padded_local_rows(h, world, BAND_ROWS)
        planes = [D.device_tensor(p, (h, w), dt, local_rank)[:lr] for p, dt in zip(view.frame_pointers(), ("<i4", "<i4", "<f4"))]
        stream = torch.cuda.ExternalStream(view.cuda_stream(), device=local_rank)

        def gather():
            with torch.cuda.stream(stream):
                return [D.gather_bands(p, h, world, BAND_ROWS) for p in planes]
    def owner_of_row(r):  # tiles_fused: the gather's rotated interleave (kernels.cuh: band_rotate); tiles_nccl: the plain one
        b = r // BAND_ROWS
        return (b % world + b // world) % world if mode == "tiles_fused" else b % world

    rays_per_rank_step = sum(1 for r in range(h) if owner_of_row(r) == rank) * w if tiles else rays_per_frame

    def barrier():
        view.synchronize()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-timed region: L2 flushed before every step, CUDA events around each step on its stream -----------------
    # tiles_fused: rank 0's event pair spans its viewport kernel (which publishes `go` to the peers) and the kernel that
    # waits for every peer's rows, i.e. the complete frame - every peer's kernel runs inside that interval (it starts after
    # `go`, the interval ends after its `done`), so rank 0's time IS the max over ranks. The peers are not timed in these
    # steps: an event between a peer's go-wait and its viewport kernel sits on the frame's critical path
    # (profiles/r02_graph_probe.json); their own kernel times come from a separate pass below (ms_per_step_per_rank).
    untimed_peer = mode == "tiles_fused" and rank != 0
    for i in range(args.warmup):
        view.set_viewport(vps[pose_index(i)])
        view.flush_l2()
        view.render(sync=True)
        if gather:
            gather()
    barrier()
    sampler.window(True)
    launches0 = view.launch_count()
    wall0 = time.perf_counter()
    dev_ms_total = 0.0
    for i in range(args.steps):
        view.set_viewport(vps[pose_index(i)])
        view.flush_l2()
        if gather:  # the NCCL gather is part of the step: one stopwatch around kernel + collective on the same stream
            view.timer_start()
            view.render(sync=False)
            gather()
            dev_ms_total += view.timer_stop()
        elif untimed_peer:
            view.render(sync=False)
            view.synchronize()
        else:
            dev_ms_total += view.render(sync=True)["kernel_ms"]
    wall1 = time.perf_counter()
    timed_launches = view.launch_count() - launches0  # every kernel this rank launched inside the timed steps (counted by the library)
    barrier()
    sampler.window(False)

    # warm-L2 variant (a viewer re-rendering the same resident tree): back-to-back frames, one event pair
    for i in range(3):
        view.render(sync=True)
    barrier()
    view.timer_start()
    for i in range(args.steps):
        view.set_viewport(vps[pose_index(i)])
        view.render(sync=False)
    warm_ms_total = view.timer_stop()
    barrier()

    # ---- the gathered frame must be the single-GPU frame, byte for byte (checked in the run, outside the timed regions) --
    gathered_equal = None
    if mode == "tiles_fused":
        view.set_viewport(vps[0])
        view.render(sync=False)
        view.synchronize()
        if rank == 0:
            got = view.read_frame()
            whole = new_view().render_to_host()
            gathered_equal = all(bool(np.array_equal(got[k].view(np.uint32), whole[k].view(np.uint32))) for k in ("hit_id", "albedo", "distance"))
        barrier()

    # ---- end to end through the public API: pose in (host), framebuffer out (pinned host planes) -----------------------
    # N = 1: view.set_viewport(pose) + view.render_to_host_async(pinned planes). N > 1 (tiles): every rank renders ITS
    # bands of the frame into its own framebuffer and copies exactly those rows into host planes all ranks share
    # (POSIX shared memory, page-locked in every process): the frame is assembled in host memory over N PCIe links.
    n_px = w * h
    shared = None
    e2e_view = view
    if tiles:
        e2e_view = new_view()
        e2e_view.set_shard(rank, world, BAND_ROWS)
        shared = D.SharedPinnedPlanes(f"{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}", n_px, 2, rank)
        ptr_sets = shared.sets
        pinned = shared.pinned
    else:
        try:
            import torch as _t  # plumbing only: pinned host buffers

            bufs = [[_t.empty(n_px, dtype=_t.int32).pin_memory(), _t.empty(n_px, dtype=_t.int32).pin_memory(),
                     _t.empty(n_px, dtype=_t.float32).pin_memory()] for _ in range(2)]
            ptr_sets = [[b.data_ptr() for b in bs] for bs in bufs]
            pinned = True
        except Exception:
            bufs = [[np.empty(n_px, dtype=np.uint32), np.empty(n_px, dtype=np.uint32), np.empty(n_px, dtype=np.float32)] for _ in range(2)]
            ptr_sets = [[a.ctypes.data for a in bs] for bs in bufs]
            pinned = False

    def e2e_loop(planes_of, pipelined=True):
        """wall ms of `steps` frames: pose in, host planes out; planes_of(set) -> the three host pointers (0 = not delivered)"""
        for i in range(4):  # warm-up through the SAME path as the timed frames (the pipelined one allocates its second slot once)
            e2e_view.set_viewport(vps[pose_index(i)])
            if pipelined:
                e2e_view.render_to_host_async_ptr(*planes_of(ptr_sets[i & 1]))
                e2e_view.wait_host(1)
            else:
                e2e_view.render_to_host_ptr(*planes_of(ptr_sets[i & 1]))
        if pipelined:
            e2e_view.wait_host(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_view.set_viewport(vps[pose_index(i)])  # the step's input: one 40-byte pose, handed over as launch parameters
            if pipelined:
                # two host plane sets, frame i's copies overlap frame i+1's kernel; every step still ends with the PREVIOUS
                # step's frame complete in host memory, the last one after the loop
                e2e_view.render_to_host_async_ptr(*planes_of(ptr_sets[i & 1]))
                e2e_view.wait_host(1)
            else:
                e2e_view.render_to_host_ptr(*planes_of(ptr_sets[0]))  # kernel + device->host copies, synchronised
        if pipelined:
            e2e_view.wait_host(0)
        if dist is not None:
            dist.barrier()  # the frame is complete when every rank's rows are in the shared planes
        return (time.perf_counter() - t0) * 1e3

    sampler.window(True)
    e2e_ms_total = e2e_loop(lambda p: p)
    e2e_sync_ms_total = e2e_loop(lambda p: p, pipelined=False)
    e2e_8b_ms_total = e2e_loop(lambda p: (p[0], 0, p[2]))  # hit id + distance: albedo is palette[hit_id & 0xFFFF]
    sampler.window(False)
    # the 8-byte frame loses nothing: the albedo plane it leaves out is the tree's colour palette looked up by hit id
    palette_equal = None
    if rank == 0 and shared is None:
        e2e_view.render_to_host_ptr(*ptr_sets[0])
        ids, alb = (np.array(bufs[0][k].numpy() if pinned else bufs[0][k], copy=True).view(np.uint32) for k in (0, 1))
        colors = tree.color_palette().astype(np.uint32)
        packed = colors[:, 0] | (colors[:, 1] << 8) | (colors[:, 2] << 16) | (colors[:, 3] << 24)
        index = ids & 0xFFFF
        has_colour = (ids != 0xFFFFFFFF) & (index != 0xFFFF)
        palette_equal = bool(np.array_equal(np.where(has_colour, packed[np.minimum(index, len(packed) - 1)], 0).astype(np.uint32), alb))
    e2e_equal = None
    if shared is not None:  # the host-assembled frame of the last full-planes loop against rank 0's own whole frame
        e2e_loop(lambda p: p, pipelined=False)
        if rank == 0:
            whole = new_view()
            whole.set_viewport(vps[pose_index(args.steps - 1)])
            wf = whole.render_to_host()
            e2e_equal = all(bool(np.array_equal(shared.plane(0, k, np.uint32), wf[name].reshape(-1).view(np.uint32)))
                            for k, name in enumerate(("hit_id", "albedo", "distance")))
        barrier()
    clocks = sampler.stop()

    per_rank_ms = [dev_ms_total / args.steps]
    if dist is not None:
        mine_ms = dev_ms_total / args.steps
        if mode == "tiles_fused":  # diagnostic pass: every member's own event pair (rank 0: the frame; a peer: its viewport kernel)
            n_diag = min(args.steps, 20)
            barrier()
            mine_ms = 0.0
            for i in range(n_diag):
                view.set_viewport(vps[pose_index(i)])
                view.flush_l2()
                mine_ms += view.render(sync=True)["kernel_ms"] / n_diag
            barrier()
        mine = torch.tensor([mine_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank_ms = [round(float(x.item()), 5) for x in every]
    if dist is not None:  # max over ranks
        t = torch.tensor([dev_ms_total, warm_ms_total, e2e_ms_total, e2e_sync_ms_total, e2e_8b_ms_total], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_total, warm_ms_total, e2e_ms_total, e2e_sync_ms_total, e2e_8b_ms_total = [float(v) for v in t.tolist()]

    frames_per_step = world if mode == "poses" else 1
    total_rays = rays_per_frame * frames_per_step * args.steps
    value = total_rays / (dev_ms_total * 1e-3) / 1e6
    value_warm = total_rays / (warm_ms_total * 1e-3) / 1e6
    st = host.stats()
    par = {"single": "one GPU renders the whole frame",
           "poses": "pose-sharded x%d (tree replicated, no exchange)" % world,
           "tiles_nccl": "one frame in interleaved bands of %d rows over %d GPUs + NCCL all_gather of compact bands" % (BAND_ROWS, world),
           "tiles_fused": "one frame in interleaved bands of %d rows over %d GPUs (tree replicated); fused gather: the traversal kernels of ranks "
                          "1..%d store into rank 0's framebuffer over NVLink (CUDA IPC), %d B/pixel on the wire, device-side go/done flags, "
                          "no host barrier and no collective per frame" % (BAND_ROWS, world, world - 1, wire_bytes)}[mode]
    e2e_frames = (world if mode == "poses" else 1) * args.steps
    # the box's device->host ceiling (plain pinned cudaMemcpyAsync, tools/d2h_ceiling.py on this pool's 8-GPU box): one GPU
    # alone, and all 8 at once (the host fabric, not the GPUs' links, limits the aggregate)
    d2h = {}
    try:
        ceil = json.load(open(ROOT / "profiles" / "r02_d2h_ceiling.json"))
        gbs = ceil["one_gpu_gbs"] if world == 1 else min(world * ceil["one_gpu_gbs"], ceil["all_gpus_aggregate_gbs"])
        e2e_gbs = 8 * n_px * e2e_frames / (e2e_8b_ms_total * 1e-3) / 1e9
        d2h = {"d2h_gbs": e2e_gbs, "d2h_ceiling_gbs": gbs, "frac_of_d2h_ceiling": e2e_gbs / gbs,
               "d2h_ceiling_source": "profiles/r02_d2h_ceiling.json: one GPU %.1f GB/s, 8 GPUs at once %.1f GB/s in total; "
                                     "N GPUs = min(N x one, the 8-GPU total)" % (ceil["one_gpu_gbs"], ceil["all_gpus_aggregate_gbs"])}
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak" if mode == "poses" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": args.workload, "description": desc, "resolution": list(res), "rays_per_step": rays_per_frame * frames_per_step,
            "rays_per_step_per_gpu": rays_per_rank_step, "poses": len(vps), "parallelism": par,
            "l2": "flushed before every timed step (384 MiB memset on the launch stream, outside the event pair)",
            "tree_bytes": st["total_bytes"], "tree_nodes": st["nodes"], "tree_bricks": st["bricks"], "tree_depth": st["depth"],
            "tree_build_s": round(t_build, 3), "voxels_inserted": int(len(scene.xyz)),
            **({"gathered_frame_equals_single_gpu_frame": gathered_equal} if mode == "tiles_fused" else {}),
            **({"mips": {"strategy": "MIPMapStrategy::default(), enabled after construction (one recalculate_mips)",
                         "viewing_distance": vd, "recalculate_s": round(t_mips, 3)}} if args.mips is not None else {}),
        },
        "ms_per_step_per_rank": per_rank_ms,
        "value_warm_l2": value_warm, "ms_per_step_warm_l2": warm_ms_total / args.steps,
        "wall_ms_per_step_incl_flush": (wall1 - wall0) * 1e3 / args.steps,
        # headline: the 8 B/pixel call; the three-plane call of the same loop is reported beside it
        "e2e": {"value": rays_per_frame * e2e_frames / (e2e_8b_ms_total * 1e-3) / 1e6, "unit": UNIT,
                "h2d_bytes_per_step": 40 * world, "d2h_bytes_per_step": 8 * n_px * (world if mode == "poses" else 1),
                "ms_per_step": e2e_8b_ms_total / args.steps,
                "planes": "hit_id + distance (svx_view_render_to_host_async with albedo = NULL): 8 B/pixel. The voxel a ray hit is its "
                          "hit id (colour index | data index << 16, the reference's PaletteIndexValues); albedo = "
                          "svx_octree_color_palette()[hit_id & 0xFFFF], a host lookup like the reference shader's color_palette[]",
                **({"albedo_from_palette_equals_albedo_plane": palette_equal} if palette_equal is not None else {}),
                **d2h,
                "three_planes": {"value": rays_per_frame * e2e_frames / (e2e_ms_total * 1e-3) / 1e6, "ms_per_step": e2e_ms_total / args.steps,
                                 "d2h_bytes_per_step": 12 * n_px * (world if mode == "poses" else 1),
                                 "synchronised_ms_per_step": e2e_sync_ms_total / args.steps,
                                 "synchronised_value": rays_per_frame * e2e_frames / (e2e_sync_ms_total * 1e-3) / 1e6,
                                 "note": "hit_id, albedo and distance planes: 12 B/pixel, the albedo plane resolved on the GPU"},
                "host_planes": ("POSIX shared memory mapped by every rank, " if shared is not None else "") + ("page-locked" if pinned else "PAGEABLE (pinning failed)"),
                **({"host_assembled_frame_equals_single_gpu_frame": e2e_equal} if shared is not None else {}),
                "note": ("per step and rank: view.set_viewport(pose) + view.render_to_host_async(shared host planes) on the rank's shard: its "
                         "kernel renders its bands, strided device->host copies put exactly those rows into the common frame, each GPU over its own "
                         "PCIe link; " if tiles else
                         "per step: view.set_viewport(pose) + view.render_to_host_async(pinned host planes) + wait for the previous frame; ")
                        + "two host plane sets, copies on a second stream overlap the next kernel; wall clock over all steps incl. the final drain"
                          " (and a barrier over the ranks). synchronised_* = the same with render_to_host and a stream sync every step"},
        "gpu_launches": int(timed_launches),
        "gpu_launches_total_incl_warmup_and_e2e": int(view.launch_count() + (e2e_view.launch_count() if e2e_view is not view else 0)),
        "clocks": clocks,
    }
    if shared is not None:
        shared.close()

    if rank == 0 and not args.no_cpu_baseline:
        chk = new_view()
        leg = oracle_leg(args.workload, scene, cams[0], res, vd, args, dev_ms_total / args.steps, rays_per_rank_step, st["total_bytes"],
                         chk.render_to_host, mips_tree=tree)
        if tiles:
            leg["roofline"]["note"] += (". N > 1: the launch is rank 0's share of the frame and its duration the whole step (viewport kernel + "
                                        "wait for the slowest peer), so this frac is a lower bound for the kernel itself")
        line.update(leg)
    if not args.no_extra and args.mips is None:
        extra = {}
        if world == 1:
            for name in ("dot_cube_1080p", "minecraft_4k", "terrain_poses_1080p"):
                if name == args.workload:
                    continue
                try:
                    extra[name] = extra_single_gpu(name, args, local_rank)
                except Exception as e:  # noqa: BLE001  an extra workload must never cost the headline line
                    extra[name] = {"error": repr(e)}
        elif tiles:
            try:
                name, out = extra_pose_batch(args, rank, local_rank, world, dist, torch)
                extra[name] = out
            except Exception as e:  # noqa: BLE001
                extra["terrain_poses_1080p"] = {"error": repr(e)}
        line["extra_workloads"] = extra
    if rank == 0:
        print(json.dumps(line))
    if mode == "tiles_fused":
        if dist is not None:
            dist.barrier()
        if rank != 0:
            view.gather_close()
    if dist is not None:
        dist.barrier()
        if rank == 0 and mode == "tiles_fused":
            view.gather_close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
