"""The lane-refill schedule (svx_view_set_schedule(view, 2); kernels.cu: render_refill_body over traverse_refill.cuh) renders
the same bits as the default schedule: it is a different assignment of pixels to lanes over time, not a different traversal."""
import numpy as np
import pytest

import shocovox_b200 as S
from shocovox_b200 import scenes

pytestmark = pytest.mark.gpu

CASES = {
    "cpu_render_64_8": (lambda: scenes.cpu_render_scene(64, 8), (640, 363)),       # brick-8 instantiation, ragged frame edges
    "cpu_render_32_2": (lambda: scenes.cpu_render_scene(32, 2), (333, 251)),       # generic code
    "dot_cube_128_32": (lambda: scenes.dot_cube_scene(128, 32), (1920, 1080)),     # brick-32, mostly sky
    "terrain_256_8_shell": (lambda: scenes.terrain_scene(256, 8, 4321, 1, shell=3), (1280, 720)),  # long crawls
}


def frames(view):
    f = view.render_to_host()
    return [f[k].view(np.uint32).copy() for k in ("hit_id", "albedo", "distance")]


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("lod", [False, True], ids=["plain", "lod"])
def test_lane_refill_frames_equal_the_static_schedules(name, lod, monkeypatch):
    make, res = CASES[name]
    scene = make()
    tree = scenes.build_tree(scene, S.Octree)
    if lod:
        tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
    host = S.OctreeGPUHost(tree, 0)
    cam = scenes.cpu_render_camera(scene.tree_size)
    vp = S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov)
    view = host.create_new_view(64, vp, res)
    if lod:
        view.set_viewing_distance(float(scene.tree_size))
    want = frames(view)
    assert (want[0] != 0xFFFFFFFF).sum() > 500
    for steps, idle, unit in ((24, 8, 8), (1, 1, 1), (3, 32, 2), (64, 16, 4)):
        monkeypatch.setenv("SVX_REFILL_STEPS", str(steps))
        monkeypatch.setenv("SVX_REFILL_MIN_IDLE", str(idle))
        monkeypatch.setenv("SVX_REFILL_UNIT", str(unit))
        other = host.create_new_view(64, vp, res)
        if lod:
            other.set_viewing_distance(float(scene.tree_size))
        other.set_schedule(2)
        for _ in range(2):  # twice: the ticket counters alternate
            got = frames(other)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
        # and a shard of the frame
        other.set_shard(1, 3, 8)
        got = frames(other)
        rows = np.array([r for r in range(res[1]) if (r // 8) % 3 == 1])
        for a, b in zip(got, want):
            assert np.array_equal(a[rows], b[rows])
