"""Level of detail on the GPU: `Octree::get_by_ray_at_lod` (src/raytracing/raytracing_on_cpu.rs:325-565) over trees whose
MIP maps are enabled (src/octree/mipmap.rs), through the C ABI, against the CPU oracle - every field bit-exact.

The reference holds no test of get_by_ray_at_lod with a finite viewing distance, so the oracle's LOD branch is pinned
only by its line-by-line restatement (MIP *construction* is pinned by the reference's mipmap KATs, tests/test_mipmap.py).
Where the reference would bounds-panic on a 4x4x4 bitmap index (after a MIP miss left the point outside the node,
:433-436, :511-514) oracle and kernel clamp identically; those rays are compared too.

All tests here need a B200 (`-m gpu`). Nothing reads /root/reference at run time.
"""
import numpy as np
import pytest

import oracle_lib as O
import shocovox_b200 as S
from shocovox_b200 import scenes
from test_gpu_parity import assert_frames_equal, bits, oracle_camera, random_rays, viewport

pytestmark = pytest.mark.gpu

F32_MAX = S.F32_MAX
BOX, POINT, POINT_BD, POSTERIZE, POSTERIZE_BD = 0, 1, 2, 3, 4


def mip_trees(scene, methods=None, thresholds=None):
    """Product and oracle trees of a scene with MIP maps switched on AFTER construction (one recalculation)."""
    tree, otree = scenes.build_tree(scene, S.Octree), scenes.build_tree(scene, O.OracleOctree)
    su = tree.albedo_mip_map_resampling_strategy()
    for lvl, (m, thr) in (methods or {}).items():
        su.set_method_at(lvl, m, thr)
        otree.set_method_at(lvl, m, thr)
    for lvl, thr in (thresholds or {}).items():
        su.set_color_similarity_thr_at(lvl, thr)
        otree.set_color_similarity_thr_at(lvl, thr)
    su.switch_albedo_mip_maps(True)
    otree.switch_albedo_mip_maps(True)
    assert tree.structure_hash() == otree.structure_hash()
    assert su.mip_hash() == otree.mip_hash()
    return tree, otree


def assert_rays_equal(g, o):
    assert np.array_equal(g["hit"], o["hit"])
    assert np.array_equal(g["palette_value"], o["palette_value"])
    assert np.array_equal(g["entry_kind"], o["entry_kind"])
    assert np.array_equal(g["rgba"], o["rgba"]) and np.array_equal(g["data"], o["data"])
    assert np.array_equal(bits(g["impact_point"]), bits(o["impact_point"]))
    assert np.array_equal(bits(g["normal"]), bits(o["normal"]))
    assert np.array_equal(bits(g["distance"]), bits(o["distance"]))


LOD_SCENES = {
    "cpu_render_64_8": lambda: scenes.cpu_render_scene(64, 8),
    "cpu_render_32_1": lambda: scenes.cpu_render_scene(32, 1),
    "cpu_render_32_2": lambda: scenes.cpu_render_scene(32, 2),
    "dot_cube_128_32": lambda: scenes.dot_cube_scene(128, 32),
    "colonnade_256_8": scenes.colonnade_scene,
    "terrain_256_8_shell": lambda: scenes.terrain_scene(256, 8, 4321, 1, shell=3),  # deeper than the 4-entry ring stack
}


@pytest.mark.parametrize("name", list(LOD_SCENES))
def test_random_rays_at_lod_bit_exact(name):
    scene = LOD_SCENES[name]()
    tree, otree = mip_trees(scene)
    host = S.OctreeGPUHost(tree)
    rays = random_rays(scene.tree_size, 20000, 77 + len(name))
    probes = 0
    for vd in (F32_MAX, 400.0, 64.0, 9.5, 1.0, 0.0, -3.0, float("inf"), float("nan")):
        g = host.get_by_rays(rays, vd)
        o = otree.get_by_rays_at_lod(rays, vd)
        assert_rays_equal(g, o)
        probes += int(o["mip_probes"].sum())
    assert probes > 10000  # the LOD branch really ran
    # and it really changes what is seen
    far, near = otree.get_by_rays_at_lod(rays, F32_MAX), otree.get_by_rays_at_lod(rays, 9.5)
    assert (far["palette_value"] != near["palette_value"]).sum() > 100


@pytest.mark.parametrize("methods", [
    {1: (BOX, 0.0), 2: (BOX, 0.0)},
    {1: (POINT_BD, 0.0), 2: (POINT_BD, 0.0), 3: (POINT_BD, 0.0)},
    {1: (POSTERIZE, 0.15), 2: (POSTERIZE_BD, 0.3), 3: (POINT, 0.0)},
], ids=["box", "point_bd", "posterize"])
def test_every_resampling_method_renders_like_the_oracle(methods):
    scene = scenes.cpu_render_scene(64, 4)
    tree, otree = mip_trees(scene, methods, {2: 0.0, 3: 0.2})
    cam = scenes.cpu_render_camera()
    view = S.OctreeGPUHost(tree).create_new_view(1, viewport(cam), (200, 150))
    for vd in (150.0, 40.0):
        view.set_viewing_distance(vd)
        assert view.viewing_distance() == vd
        ora = otree.render(oracle_camera(cam), 200, 150, viewing_distance=vd)
        assert_frames_equal(view.render_to_host(), ora)
        assert ora["mip_probes"] > 0


def test_frames_at_lod_c1_and_deep_terrain():
    """Whole frames: the cpu_render scene and a tree deeper than the ring stack (mip_level drifts there, see
    traverse.cuh), static and persistent schedules, at the viewing distance the reference's GPU path would use
    (viewport.frustum.z) and at closer ones."""
    for scene, cam, res in [
        (scenes.cpu_render_scene(), scenes.cpu_render_camera(), (320, 240)),
        (scenes.terrain_scene(256, 8, 4321, 1, shell=4), scenes.terrain_camera(256), (320, 180)),
        (scenes.colonnade_scene(), scenes.colonnade_camera(), (320, 180)),
    ]:
        tree, otree = mip_trees(scene)
        host = S.OctreeGPUHost(tree)
        view = host.create_new_view(1, viewport(cam), res)
        assert view.viewing_distance() == pytest.approx(F32_MAX)
        for vd in (F32_MAX, 500.0, 120.0, float(cam.frustum[2]), 2.0):
            view.set_viewing_distance(vd)
            ora = otree.render(oracle_camera(cam), res[0], res[1], viewing_distance=vd)
            for persistent in (False, True):
                view.set_schedule(persistent)
                assert_frames_equal(view.render_to_host(), ora)


def test_mips_follow_edits_and_reload():
    """insert / clear refresh the MIPs incrementally (insert.rs:371, clear.rs:335); reload ships the touched MIP bricks."""
    scene = scenes.cpu_render_scene()
    tree, otree = mip_trees(scene)
    cam = scenes.cpu_render_camera()
    host = S.OctreeGPUHost(tree)
    view = host.create_new_view(1, viewport(cam), (200, 150))
    view.set_viewing_distance(90.0)
    assert_frames_equal(view.render_to_host(), otree.render(oracle_camera(cam), 200, 150, viewing_distance=90.0))
    rng = np.random.default_rng(3)
    for _ in range(200):
        p = tuple(int(v) for v in rng.integers(0, 64, 3))
        if rng.random() < 0.7:
            c = int(rng.integers(1, 2**32 - 1)) | 0xFF
            tree.insert(p, c)
            otree.insert(p, c)
        else:
            tree.clear(p)
            otree.clear(p)
    tree.insert_at_lod((32, 32, 0), 16, 0x33CC33FF)
    otree.insert_at_lod((32, 32, 0), 16, 0x33CC33FF)
    assert tree.albedo_mip_map_resampling_strategy().mip_hash() == otree.mip_hash()
    view.reload()
    up = host.last_upload()
    assert not up["full"] and 0 < up["bricks"] < host.stats()["bricks"]
    assert_frames_equal(view.render_to_host(), otree.render(oracle_camera(cam), 200, 150, viewing_distance=90.0))


def test_switching_mips_off_restores_get_by_ray():
    scene = scenes.cpu_render_scene()
    tree, otree = mip_trees(scene)
    plain = scenes.build_tree(scene, O.OracleOctree)
    cam = scenes.cpu_render_camera()
    host = S.OctreeGPUHost(tree)
    view = host.create_new_view(1, viewport(cam), (160, 120))
    view.set_viewing_distance(30.0)
    with_mips = view.render_to_host()
    tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(False)
    view.reload()
    without = view.render_to_host()
    # the MIP pass only added palette colours: hit ids of the plain tree are unchanged
    assert_frames_equal(without, plain.render(oracle_camera(cam), 160, 120))
    assert (with_mips["hit_id"] != without["hit_id"]).any()


def test_single_ray_get_by_ray_at_lod_mirror():
    """`tree.get_by_ray_at_lod(&ray, viewing_distance)` through the Python mirror; MIP hits carry no user data."""
    t = S.Octree(16, 4)
    o = O.OracleOctree(16, 4)
    for tr in (t, o):
        for x in range(4):
            for z in range(4):
                tr.insert((x, 0, z), 0xFF0000FF if (x + z) % 2 else 0x00FF00FF, 7)
    t.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True).set_method_at(1, S.MIP_BOX_FILTER).recalculate_mips()
    o.switch_albedo_mip_maps(True).set_method_at(1, BOX).recalculate_mips()
    ray = S.Ray((2.2, 20.0, 1.7), (0.0, -1.0, 0.0))
    near = t.get_by_ray(ray)
    far = t.get_by_ray_at_lod(ray, 4.0)
    on, of = o.get_by_ray((2.2, 20.0, 1.7), (0.0, -1.0, 0.0)), o.get_by_ray_at_lod((2.2, 20.0, 1.7), (0.0, -1.0, 0.0), 4.0)
    assert near is not None and far is not None and on.hit and of.hit
    assert near.palette_value == on.palette_value and far.palette_value == of.palette_value
    assert near.entry.data == 7 and far.entry.data is None  # "Simplified views do not contain user data!" (:323)
    assert far.entry.albedo not in (S.Albedo.from_u32(0xFF0000FF), S.Albedo.from_u32(0x00FF00FF))  # a blended colour
    assert far.impact_point == tuple(float(v) for v in of.impact_point)


def test_crawl_under_lod_is_exact_on_a_large_tree():
    """With MIP maps on, every failing root iteration of the 0.1-nudge crawl also raises mip_level and re-evaluates the
    LOD test; the kernel fast-forwards only once lod_quiescent() (traverse.cuh) proves the test stays false. Rays
    skimming a 512^3 terrain, from outside and from inside, over viewing distances on both sides of that switch."""
    scene = scenes.terrain_scene(512, 8, 4321, 1, shell=4)
    tree, otree = mip_trees(scene)
    cam = scenes.terrain_camera(512)
    ocam = oracle_camera(cam)
    skim = np.stack([O.pixel_ray(ocam, 320, 180, x, y) for y in range(0, 180, 5) for x in range(0, 320, 5)])
    rng = np.random.default_rng(9)
    origin = rng.uniform(1, 511, (2000, 3)).astype(np.float32)
    origin[:, 1] = rng.uniform(200, 511, 2000).astype(np.float32)
    d = rng.normal(size=(2000, 3)).astype(np.float32)
    d[:, 1] = np.abs(d[:, 1]) * 0.2
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    inside = np.concatenate([origin, (d / ln[:, None]).astype(np.float32)], axis=1)
    host = S.OctreeGPUHost(tree)
    crawled = 0
    for rays in (skim, inside):
        for vd in (F32_MAX, 1.0e6, 5000.0, 700.0, 90.0, 12.0, 4.0, 3.9, 0.7):
            g = host.get_by_rays(rays, vd)
            o = otree.get_by_rays_at_lod(rays, vd)
            assert_rays_equal(g, o)
            crawled = max(crawled, int(o["outer_iters"].max()))
    assert crawled > 2000  # the crawl really happens
