#!/usr/bin/env python3
"""Primary-ray throughput of the B200 path on the reference's named configs (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step = one pass of the hot path over one batch: one full frame (in-kernel ray generation + get_by_ray per pixel +
framebuffer write) per GPU. Default workload = BASELINE configs[1], examples/dot_cube at 1920x1080 on one B200.
N > 1 (torchrun, one rank per GPU): camera poses are sharded over the ranks (batch mode of the north star), the tree is
replicated, there is no data-path collective -> weak scaling. `--mode tiles` instead splits ONE frame into row bands and
gathers the bands with NCCL (strong scaling).

Prints ONE JSON line (rank 0). `value` is device-timed with CUDA events on the launching stream, L2 flushed before
every timed step; `e2e` goes through the public API with host buffers (pose in, framebuffer out, copies timed).
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

WORKLOADS = {
    # name: (scene factory, camera factory, (w, h), description)
    "dot_cube_1080p": ("dot_cube", dict(zoom=True), (1920, 1080),
                       "examples/dot_cube.rs tree 256/32, camera (512,128,-512)->0, glass 10x10 at frustum.z=200 "
                       "(the example's own CPU ray loop, dot_cube.rs:204-231), 1920x1080"),
    "dot_cube_1080p_fov": ("dot_cube", dict(zoom=False), (1920, 1080),
                           "examples/dot_cube.rs tree 256/32, glass 10x10 at fov=3 (shader placement), 1920x1080"),
    "dot_cube_4k": ("dot_cube", dict(zoom=True), (3840, 2160), "examples/dot_cube.rs tree 256/32, glass at frustum.z, 3840x2160"),
    "cpu_render_150": ("cpu_render", {}, (150, 150), "examples/cpu_render.rs tree 64/8, 150x150 (BASELINE configs[0])"),
    "cpu_render_1080p": ("cpu_render", {}, (1920, 1080), "examples/cpu_render.rs tree 64/8, 1920x1080"),
    "cpu_render_4k": ("cpu_render", {}, (3840, 2160), "examples/cpu_render.rs tree 64/8, 3840x2160"),
    "minecraft_4k": ("minecraft", {}, (3840, 2160), "synthetic minecraft-style blocky heightfield 1024/32, 3840x2160 (BASELINE configs[2])"),
}


def make_workload(name: str):
    from shocovox_b200 import scenes

    kind, cam_kw, res, desc = WORKLOADS[name]
    if kind == "dot_cube":
        return scenes.dot_cube_scene(), scenes.dot_cube_camera(**cam_kw), res, desc
    if kind == "cpu_render":
        return scenes.cpu_render_scene(), scenes.cpu_render_camera(), res, desc
    if kind == "minecraft":
        return scenes.terrain_scene(1024, 32, 1234, 4, shell=8, name="minecraft"), scenes.terrain_camera(1024), res, desc
    raise KeyError(name)


# ---- clocks -------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU from a thread while the timed region runs (NVML)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._active = threading.Event()
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
                    for bit, name in self.REASONS.items():
                        if mask & bit and name != "gpu_idle":
                            self.reasons.add(name)
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
def oracle_frame_stats(scene, cam, res, min_seconds: float, max_frames: int):
    """Times the CPU oracle (all host threads) on whole frames of the workload; returns throughput + per-ray counts."""
    import oracle_lib as O
    from shocovox_b200 import scenes

    tree = scenes.build_tree(scene, O.OracleOctree)
    ocam = O.make_camera(cam.origin, cam.direction, cam.frustum[0], cam.frustum[1], cam.glass_distance)
    threads = int(O.lib().svxo_hardware_threads())
    times, last = [], None
    t_total = 0.0
    while len(times) < max_frames and (t_total < min_seconds or len(times) < 2):
        last = tree.render(ocam, res[0], res[1], threads=threads)
        times.append(last["seconds"])
        t_total += last["seconds"]
    rays = res[0] * res[1]
    return {"tree": tree, "frame": last, "threads": threads, "times": times, "rays": rays,
            "mrays_best": rays / min(times) / 1e6, "mrays_mean": rays / (sum(times) / len(times)) / 1e6}


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU get_by_ray loop (the oracle port: the Rust crate cannot be built here)."""
    if rank != 0:
        return 0
    scene, cam, res, desc = make_workload(args.workload)
    import oracle_lib as O
    from shocovox_b200 import scenes

    tree = scenes.build_tree(scene, O.OracleOctree)
    ocam = O.make_camera(cam.origin, cam.direction, cam.frustum[0], cam.frustum[1], cam.glass_distance)
    threads = int(O.lib().svxo_hardware_threads())
    for _ in range(args.warmup):
        tree.render(ocam, res[0], res[1], threads=threads)
    t = 0.0
    for _ in range(args.steps):
        t += tree.render(ocam, res[0], res[1], threads=threads)["seconds"]
    rays = res[0] * res[1]
    value = rays * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "resolution": list(res), "rays_per_step": rays},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} whole frames of the workload ({rays} rays each), rows interleaved over {threads} host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU get_by_ray of shocovox-rs restated in C++ (oracle/): the Rust crate cannot be compiled in this image (no cargo/rustc)",
    }
    print(json.dumps(line))
    return 0


# ---- our arm ---------------------------------------------------------------------------------------------------------
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dot_cube_1080p", choices=list(WORKLOADS))
    ap.add_argument("--mode", default="poses", choices=["poses", "tiles"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also measure the other named workloads (kernel time only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import shocovox_b200 as S

    if S.cuda_device_count() < 1:
        print(json.dumps({"error": "no CUDA device: the ray path has no CPU fallback"}))
        return 1
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene, cam, res, desc = make_workload(args.workload)
    from shocovox_b200 import scenes

    t_build = time.time()
    tree = scenes.build_tree(scene, S.Octree)
    host = S.OctreeGPUHost(tree, local_rank)
    t_build = time.time() - t_build
    vp = S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov)
    view = host.create_new_view(64, vp, res)
    if cam.glass_at_frustum_z:
        view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
    tiles = args.mode == "tiles" and world > 1
    if tiles:
        view.set_shard(rank, world, 8)
    rays_per_frame = res[0] * res[1]
    rays_per_rank_step = rays_per_frame // world if tiles else rays_per_frame

    def barrier():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        view.synchronize()

    gather_buf = None
    if tiles:
        # framebuffer bands travel to every rank with one NCCL all_gather per plane set (torch is plumbing here)
        gather_buf = None  # set up lazily below

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-timed region: L2 flushed before every step, CUDA events around each render on its stream ----------
    for _ in range(args.warmup):
        view.flush_l2()
        view.render(sync=True)
    barrier()
    sampler.window(True)
    wall0 = time.perf_counter()
    kernel_ms = []
    for _ in range(args.steps):
        view.flush_l2()
        kernel_ms.append(view.render(sync=True)["kernel_ms"])
    wall1 = time.perf_counter()
    barrier()
    sampler.window(False)
    dev_ms_total = float(sum(kernel_ms))

    # warm-L2 variant (a viewer re-rendering the same resident tree): back-to-back launches, one event pair
    for _ in range(3):
        view.render(sync=True)
    view.timer_start()
    for _ in range(args.steps):
        view.render(sync=False)
    warm_ms_total = view.timer_stop()

    # ---- end to end through the public API: pose in (host), framebuffer out (pinned host) --------------------------
    import ctypes

    n_px = res[0] * res[1]
    pinned = []
    if torch is None:
        try:
            import torch  # plumbing only: pinned host buffers
        except Exception:
            torch = None
    if torch is not None:
        bufs = [torch.empty(n_px, dtype=torch.int32).pin_memory() for _ in range(2)] + [torch.empty(n_px, dtype=torch.float32).pin_memory()]
        ptrs = [b.data_ptr() for b in bufs]
        pinned = bufs
    else:
        arrs = [np.empty(n_px, dtype=np.uint32), np.empty(n_px, dtype=np.uint32), np.empty(n_px, dtype=np.float32)]
        ptrs = [a.ctypes.data for a in arrs]
        pinned = arrs
    for _ in range(3):
        view.set_viewport(vp)
        view.render_to_host_ptr(*ptrs)
    barrier()
    sampler.window(True)
    e0 = time.perf_counter()
    for _ in range(args.steps):
        view.set_viewport(vp)              # the step's input: one 40-byte pose, handed over as launch parameters
        view.render_to_host_ptr(*ptrs)     # kernel + three device->host copies, synchronised
    e1 = time.perf_counter()
    sampler.window(False)
    barrier()
    e2e_ms_total = (e1 - e0) * 1e3
    clocks = sampler.stop()
    launches = view.launch_count()

    # max over ranks
    if dist is not None:
        t = torch.tensor([dev_ms_total, warm_ms_total, e2e_ms_total], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_total, warm_ms_total, e2e_ms_total = [float(v) for v in t.tolist()]

    total_rays = rays_per_rank_step * world * args.steps
    value = total_rays / (dev_ms_total * 1e-3) / 1e6
    value_warm = total_rays / (warm_ms_total * 1e-3) / 1e6
    e2e_value = total_rays / (e2e_ms_total * 1e-3) / 1e6

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": args.workload, "description": desc, "resolution": list(res), "rays_per_step_per_gpu": rays_per_rank_step,
            "parallelism": ("row bands of 8 over %d GPUs + NCCL gather" % world) if tiles else ("pose-sharded x%d (tree replicated, no collective)" % world),
            "l2": "flushed before every timed step (384 MiB memset on the launch stream, outside the event pair); tree is smaller than L2",
            "tree_bytes": host.stats()["total_bytes"], "tree_nodes": host.stats()["nodes"], "tree_bricks": host.stats()["bricks"],
            "tree_depth": host.stats()["depth"], "tree_build_s": round(t_build, 3),
        },
        "value_warm_l2": value_warm, "ms_per_step_warm_l2": warm_ms_total / args.steps,
        "wall_ms_per_step_incl_flush": (wall1 - wall0) * 1e3 / args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 40, "d2h_bytes_per_step": 12 * n_px,
                "ms_per_step": e2e_ms_total / args.steps,
                "note": "view.set_viewport(pose) + view.render_to_host(pinned hit_id, albedo, distance); wall clock, synchronised every step"},
        "gpu_launches": int(args.steps),
        "gpu_launches_total_incl_warmup_and_e2e": int(launches),
        "clocks": clocks,
    }

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        traffic = None
        tpath = ROOT / "profiles" / "ncu_traffic.json"
        if tpath.exists():
            traffic = json.loads(tpath.read_text()).get(args.workload)
        if not args.no_cpu_baseline:
            o = oracle_frame_stats(scene, cam, res, min_seconds=6.0, max_frames=12)
            f = o["frame"]
            alg_bytes = 12 * o["rays"] + 16 * f["node_iters"] + 4 * f["voxel_fetches"]
            t_kernel = dev_ms_total / args.steps * 1e-3
            achieved = alg_bytes / (rays_per_frame / rays_per_rank_step) / t_kernel / 1e9
            line["roofline"] = {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": "svx::render_kernel",
                "algorithmic_bytes_per_launch": alg_bytes,
                "per_ray": {"node_visits": f["node_iters"] / o["rays"], "voxel_fetches": f["voxel_fetches"] / o["rays"],
                            "bytes": alg_bytes / o["rays"], "rays_entering_root": f["rays_in_root"] / o["rays"]},
                "compulsory_bound_ms": (host.stats()["total_bytes"] + 12 * o["rays"]) / (peak * 1e9) * 1e3,
                "note": "B_ray = 12 + 16 N_node + 4 N_vox, N counted by the CPU oracle on the same rays (SURVEY 8(d)); the path is latency-bound pointer chasing, the HBM fraction is small by construction",
            }
            line["cpu_baseline"] = {
                "value": o["mrays_best"], "unit": UNIT, "cores": o["threads"], "kind": "port",
                "sample": f"{len(o['times'])} whole frames of the workload ({o['rays']} rays each), rows interleaved over {o['threads']} host threads; best frame",
                "mean": o["mrays_mean"], "ms_per_frame_best": min(o["times"]) * 1e3,
            }
            # parity spot check of what was just timed (outside every timed region)
            got = view.render_to_host()
            line["parity_vs_oracle"] = {
                "hit_id_equal": bool(np.array_equal(got["hit_id"], f["hit_id"])) if not tiles else None,
                "distance_bits_equal": bool(np.array_equal(got["distance"].view(np.uint32), f["distance"].view(np.uint32))) if not tiles else None,
            }
        if args.extra and world == 1:
            extra = {}
            for name in ("dot_cube_1080p_fov", "cpu_render_1080p", "cpu_render_4k", "dot_cube_4k"):
                sc2, cam2, res2, _ = make_workload(name)
                tr2 = scenes.build_tree(sc2, S.Octree) if sc2.name != scene.name else tree
                h2 = S.OctreeGPUHost(tr2, local_rank)
                v2 = h2.create_new_view(64, S.Viewport(cam2.origin, cam2.direction, cam2.frustum, cam2.fov), res2)
                if cam2.glass_at_frustum_z:
                    v2.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
                ms = []
                for i in range(13):
                    v2.flush_l2()
                    k = v2.render(sync=True)["kernel_ms"]
                    if i >= 3:
                        ms.append(k)
                extra[name] = {"ms_per_frame": float(np.mean(ms)), "mrays_per_s": res2[0] * res2[1] / (np.mean(ms) * 1e-3) / 1e6}
            line["extra_workloads"] = extra
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
