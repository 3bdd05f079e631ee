"""N-GPU framebuffer == 1-GPU framebuffer, byte for byte, across REAL devices. Needs >= 2 GPUs: skipped on a
1-GPU box (the same protocol runs there on one device, tests/test_gpu_gather.py)."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import shocovox_b200 as S
from shocovox_b200 import scenes
from test_gpu_parity import bits, viewport

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def usable_world():
    n = S.cuda_device_count()
    return 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 0


@pytest.mark.parametrize("workload", ["cpu_render_1080p", "dot_cube_1080p"])
def test_tile_sharded_frame_equals_single_gpu(workload):
    """One process per GPU (torchrun), every GPU the box has (2, 4 or 8): fused gather in both wire formats + the NCCL path."""
    world = usable_world()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "multi_gpu_check.py"), "--workload", workload, "--steps", "5"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    report = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert report["world"] == world and report["all_ranks_ok"]
    for mode in ("fused_12B", "fused_8B", "nccl_gather"):
        assert report[mode]["equal_to_single_gpu"], mode


@pytest.mark.parametrize("wire", [S.WIRE_THREE_PLANES, S.WIRE_ID_DISTANCE], ids=["12B", "8B"])
def test_svx_multi_across_devices(wire):
    """One process drives all GPUs (svx_multi_*): gathered frame, host-assembled frame and a pose batch."""
    world = usable_world()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    tree = scenes.build_tree(scenes.cpu_render_scene(), S.Octree)
    cams = [scenes.cpu_render_camera(k=5 * i) for i in range(world + 1)]
    res = (1280, 720)
    single = S.OctreeGPUHost(tree, 0).create_new_view(1, viewport(cams[0]), res)
    want = []
    for c in cams:
        single.set_viewport(viewport(c))
        want.append(single.render_to_host())
    m = S.MultiGPU(tree, list(range(world)), viewport(cams[0]), res, rows_per_band=8, wire=wire)

    def same(got, ref, what):
        for k in ("hit_id", "albedo", "distance"):
            assert np.array_equal(bits(got[k]) if k == "distance" else got[k], bits(ref[k]) if k == "distance" else ref[k]), (what, k)

    for i in range(3):
        m.set_viewport(viewport(cams[i]))
        same(m.read_root_frame(), want[i], f"gathered {i}")
        same(m.render_to_host(), want[i], f"host-assembled {i}")
    batch = m.render_poses([viewport(c) for c in cams])
    for i in range(len(cams)):
        same({k: batch[k][i] for k in ("hit_id", "albedo", "distance")}, want[i], f"pose {i}")
