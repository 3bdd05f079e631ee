#!/usr/bin/env python3
"""What would a CUDA graph buy the viewport path (VERDICT r1 "next" item 2)? K frames queued back to back on a view's stream
- the way bench.py and svx_multi_render queue them, no host synchronisation in between - against the same K launches captured
into one CUDA graph and replayed. Per-frame device time (events around the whole batch), one GPU; whole sponza 4K frames,
rank 0's eighth of one (the 8-GPU share), and a small frame where the launch itself is the cost.

    python tools/graph_probe.py
"""
import json, sys, time
from pathlib import Path
import numpy as np
from cuda.bindings import runtime as rt

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import shocovox_b200 as S  # noqa: E402
from shocovox_b200 import scenes  # noqa: E402

K = 64


def ok(ret):
    err = ret[0] if isinstance(ret, tuple) else ret
    assert int(err) == 0, ret
    return ret[1] if isinstance(ret, tuple) and len(ret) > 1 else None


def probe(view):
    stream = view.cuda_stream()
    for _ in range(8):
        view.render(sync=False)
    view.synchronize()
    out = {}
    best = 1e9
    for _ in range(5):
        view.timer_start()
        t0 = time.perf_counter()
        for _ in range(K):
            view.render(sync=False)
        host_us = (time.perf_counter() - t0) * 1e6 / K
        best = min(best, view.timer_stop() * 1e3 / K)
    out["stream_us_per_frame"] = round(best, 3)
    out["host_us_per_queued_frame"] = round(host_us, 2)
    want = view.read_frame()
    ok(rt.cudaStreamBeginCapture(stream, rt.cudaStreamCaptureMode.cudaStreamCaptureModeRelaxed))
    for _ in range(K):
        view.render(sync=False)
    graph = ok(rt.cudaStreamEndCapture(stream))
    exe = ok(rt.cudaGraphInstantiate(graph, 0))
    ok(rt.cudaGraphLaunch(exe, stream))
    view.synchronize()
    best = 1e9
    for _ in range(5):
        view.timer_start()
        ok(rt.cudaGraphLaunch(exe, stream))
        best = min(best, view.timer_stop() * 1e3 / K)
    out["graph_us_per_frame"] = round(best, 3)
    got = view.read_frame()
    out["frames_equal"] = all(bool(np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32))) for k in ("hit_id", "albedo", "distance"))
    ok(rt.cudaGraphExecDestroy(exe))
    ok(rt.cudaGraphDestroy(graph))
    return out


result = {"frames_per_batch": K}
scene, cams, res, _ = bench.make_workload("sponza_4k")
tree = scenes.build_tree(scene, S.Octree)
host = S.OctreeGPUHost(tree, 0)
cam = cams[0]
for label, world in (("sponza_4k_whole", 1), ("sponza_4k_eighth", 8)):
    view = host.create_new_view(64, S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov), res)
    if world > 1:
        view.set_shard(0, world, 8)
    result[label] = probe(view)
    print(label, json.dumps(result[label]), flush=True)
    del view
small = scenes.cpu_render_scene()
shost = S.OctreeGPUHost(scenes.build_tree(small, S.Octree), 0)
c = scenes.cpu_render_camera()
view = shost.create_new_view(64, S.Viewport(c.origin, c.direction, c.frustum, c.fov), (256, 144))
result["cpu_render_256x144"] = probe(view)
print(json.dumps(result))
