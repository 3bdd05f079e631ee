"""Parity of the CUDA path (through the C ABI) with the CPU oracle: hit voxel id and albedo bit-exact, hit distance
within 1e-4 relative (BASELINE.json north star) - and, stricter, bit-exact impact points / normals / distances.

All tests here need a B200 (`-m gpu`). Nothing reads /root/reference at run time.
"""
import zlib

import numpy as np
import pytest

import oracle_lib as O
import shocovox_b200 as S
from product_adapter import ProductOctree
from ray_cases import CASES, check_expectation
from shocovox_b200 import scenes

pytestmark = pytest.mark.gpu

DIST_RTOL = 1e-4  # tolerance stated by BASELINE.json's north star


def bits(a):
    """f32 bit patterns, with every NaN mapped to one canonical pattern (0/0 gives 0xFFC00000 on x86 and 0x7FFFFFFF
    on the GPU; a NaN's sign and payload carry no meaning - the reference's normal of a hit at the exact cell centre)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.where(np.isnan(a), np.uint32(0x7FC00000), a.view(np.uint32))


def oracle_camera(c):
    return O.make_camera(c.origin, c.direction, c.frustum[0], c.frustum[1], c.glass_distance)


def viewport(c):
    return S.Viewport(c.origin, c.direction, c.frustum, c.fov)


def both_trees(scene):
    return scenes.build_tree(scene, S.Octree), scenes.build_tree(scene, O.OracleOctree)


def assert_frames_equal(gpu, ora, strict=True):
    assert np.array_equal(gpu["hit_id"], ora["hit_id"]), "hit voxel ids differ"
    ora_rgba = ora["albedo"].reshape(ora["albedo"].shape[0], ora["albedo"].shape[1], 4).copy().view(np.uint32)[..., 0]
    assert np.array_equal(gpu["albedo"], ora_rgba), "albedo differs"
    d_gpu, d_ora = gpu["distance"].astype(np.float64), ora["distance"].astype(np.float64)
    rel = np.abs(d_gpu - d_ora) / np.maximum(np.abs(d_ora), 1e-30)
    rel[d_ora == 0] = np.abs(d_gpu[d_ora == 0])
    assert rel.max() <= DIST_RTOL, f"distance rel err {rel.max()}"
    if strict:
        assert np.array_equal(bits(gpu["distance"]), bits(ora["distance"])), "distance bits differ"


def render_pair(scene, cam, res):
    tree, otree = both_trees(scene)
    assert tree.structure_hash() == otree.structure_hash()
    host = S.OctreeGPUHost(tree)
    view = host.create_new_view(64, viewport(cam), res)
    if cam.glass_at_frustum_z:
        view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
    gpu = view.render_to_host()
    ora = otree.render(oracle_camera(cam), res[0], res[1])
    assert ora["would_panic"] == 0
    return gpu, ora, view, host, tree, otree


# ---- the reference's deterministic edge-case rays, on the GPU (src/raytracing/tests.rs:253-813) ------------------
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_edge_case_rays_on_gpu(case):
    p, o = ProductOctree(case["size"], case["dim"]), O.OracleOctree(case["size"], case["dim"])
    case["build"](p)
    case["build"](o)
    ray = np.concatenate([np.asarray(case["origin"], np.float32), np.asarray(case["direction"], np.float32)])
    host = S.OctreeGPUHost(p.tree)
    g = host.get_by_rays(ray[None])[0]
    h = o.get_by_ray(case["origin"], case["direction"])
    check_expectation(case, bool(g["hit"]), int(g["entry_kind"]), tuple(int(c) for c in g["rgba"]), int(g["data"]),
                      tuple(g["normal"]))
    assert bool(g["hit"]) == bool(h.hit)
    if h.hit:
        assert int(g["palette_value"]) == h.palette_value
        assert np.array_equal(bits(g["impact_point"]), bits(np.array(h.impact_point[:], np.float32)))
        assert np.array_equal(bits(g["normal"]), bits(np.array(h.normal[:], np.float32)))
        assert bits(g["distance"]) == bits(np.float32(h.distance))


def test_single_ray_api_matches_reference_semantics():
    """`Octree::get_by_ray(&Ray)` through the mirror API (one ray, GPU)."""
    case = next(c for c in CASES if c["name"] == "detailed_brick_z_edge_error")
    p = ProductOctree(case["size"], case["dim"])
    case["build"](p)
    hit = p.tree.get_by_ray(S.Ray(case["origin"], case["direction"]))
    assert hit is not None and hit.entry == S.entry(1) and hit.normal == (0.0, 0.0, -1.0)
    miss = p.tree.get_by_ray(S.Ray((100.0, 100.0, 100.0), tuple(S.normalized((1, 1, 1)))))
    assert miss is None


# ---- random rays: every output field bit-exact --------------------------------------------------------------------
def random_rays(size, n, seed):
    rng = np.random.default_rng(seed)
    origin = rng.uniform(-1.0 * size, 2.0 * size, (n, 3)).astype(np.float32)
    inside = rng.random(n) < 0.25
    origin[inside] = rng.uniform(0, size, (int(inside.sum()), 3)).astype(np.float32)
    target = rng.uniform(0, size, (n, 3)).astype(np.float32)
    d = target - origin
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    d = (d / ln[:, None]).astype(np.float32)
    # a share of axis-parallel and grazing rays (infinite / NaN scale factors)
    k = n // 16
    d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], k)[:, None].astype(np.float32)
    origin[:k] = np.round(origin[:k])
    return np.concatenate([origin, d], axis=1)


SCENES_FOR_RAYS = {
    "cpu_render_64_8": lambda: scenes.cpu_render_scene(64, 8),
    "cpu_render_32_1": lambda: scenes.cpu_render_scene(32, 1),
    "cpu_render_32_2": lambda: scenes.cpu_render_scene(32, 2),
    "cpu_render_64_4": lambda: scenes.cpu_render_scene(64, 4),
    "dot_cube_128_32": lambda: scenes.dot_cube_scene(128, 32),
    "criterion_128_8": lambda: scenes.criterion_scene(128, 8, 48),
    "colonnade_256_8": scenes.colonnade_scene,
    "terrain_blocky_128_8": lambda: scenes.terrain_scene(128, 8, 1234, 4),
    "terrain_256_8_shell": lambda: scenes.terrain_scene(256, 8, 4321, 1, shell=3),
}


@pytest.mark.parametrize("name", list(SCENES_FOR_RAYS))
def test_random_rays_bit_exact(name):
    scene = SCENES_FOR_RAYS[name]()
    tree, otree = both_trees(scene)
    assert tree.structure_hash() == otree.structure_hash()
    rays = random_rays(scene.tree_size, 60000, zlib.crc32(name.encode()) % 1000)
    g = S.OctreeGPUHost(tree).get_by_rays(rays)
    o = otree.get_by_rays(rays)
    assert int(o["would_panic"].sum()) == 0
    assert np.array_equal(g["hit"], o["hit"])
    assert np.array_equal(g["palette_value"], o["palette_value"])
    assert np.array_equal(g["entry_kind"], o["entry_kind"])
    assert np.array_equal(g["rgba"], o["rgba"]) and np.array_equal(g["data"], o["data"])
    assert np.array_equal(bits(g["impact_point"]), bits(o["impact_point"]))
    assert np.array_equal(bits(g["normal"]), bits(o["normal"]))
    assert np.array_equal(bits(g["distance"]), bits(o["distance"]))
    assert o["hit"].sum() > 1000  # the test actually exercises hits


# ---- frames: the BASELINE configs --------------------------------------------------------------------------------------
def test_c1_cpu_render_150x150_sequence():
    """BASELINE config 1: examples/cpu_render.rs scene at 150x150, fixed camera sequence angle_k = 40 + 0.005 k."""
    scene = scenes.cpu_render_scene()
    tree, otree = both_trees(scene)
    host = S.OctreeGPUHost(tree)
    cams = [scenes.cpu_render_camera(64, k) for k in range(0, 64, 7)]
    view = host.create_new_view(64, viewport(cams[0]), (150, 150))
    batch = view.render_batch([viewport(c) for c in cams])
    for i, c in enumerate(cams):
        ora = otree.render(oracle_camera(c), 150, 150)
        assert_frames_equal({k: batch[k][i] for k in ("hit_id", "albedo", "distance")}, ora)
        assert (ora["hit_id"] != S.MISS).sum() > 1000


def test_c1_cpu_render_1080p():
    gpu, ora, *_ = render_pair(scenes.cpu_render_scene(), scenes.cpu_render_camera(), (1920, 1080))
    assert_frames_equal(gpu, ora)


@pytest.mark.parametrize("zoom", [False, True], ids=["glass_at_fov", "glass_at_frustum_z"])
def test_c2_dot_cube_1080p(zoom):
    """BASELINE config 2: examples/dot_cube.rs tree (256 / 32) at 1920x1080, both looking-glass placements."""
    gpu, ora, *_ = render_pair(scenes.dot_cube_scene(), scenes.dot_cube_camera(zoom=zoom), (1920, 1080))
    assert_frames_equal(gpu, ora)
    assert (ora["hit_id"] != S.MISS).sum() > 10000


def test_c3_blocky_terrain_reduced():
    """BASELINE config 3 at a size the oracle finishes in seconds (256^3 / 32, 1280x720)."""
    gpu, ora, *_ = render_pair(scenes.terrain_scene(256, 32, 1234, 4, name="minecraft"), scenes.terrain_camera(256), (1280, 720))
    assert_frames_equal(gpu, ora)
    assert (ora["hit_id"] != S.MISS).sum() > 100000


def test_c4_mixed_resolution_bricks():
    """BASELINE config 4's tree family: insert_at_lod slabs (Solid bricks, UniformLeaf nodes) + per-voxel detail."""
    gpu, ora, *_ = render_pair(scenes.colonnade_scene(), scenes.colonnade_camera(), (1280, 720))
    assert_frames_equal(gpu, ora)
    assert (ora["hit_id"] != S.MISS).sum() > 100000


def test_c5_deep_tree_poses():
    """BASELINE config 5's tree family (brick 8, deeper than the 4-entry ring stack) over orbiting poses."""
    scene = scenes.terrain_scene(256, 8, 4321, 1, shell=4)
    tree, otree = both_trees(scene)
    host = S.OctreeGPUHost(tree)
    assert host.stats()["depth"] > 4
    cams = scenes.orbit_cameras(256, 8)
    view = host.create_new_view(64, viewport(cams[0]), (640, 360))
    batch = view.render_batch([viewport(c) for c in cams])
    for i, c in enumerate(cams):
        ora = otree.render(oracle_camera(c), 640, 360)
        assert_frames_equal({k: batch[k][i] for k in ("hit_id", "albedo", "distance")}, ora)


# ---- edge cases ------------------------------------------------------------------------------------------------------
def test_empty_tree_renders_all_misses():
    t = S.Octree(64, 8)
    view = S.OctreeGPUHost(t).create_new_view(1, viewport(scenes.cpu_render_camera()), (64, 48))
    f = view.render_to_host()
    assert (f["hit_id"] == S.MISS).all() and (f["albedo"] == 0).all() and (f["distance"] == 0).all()


def test_ragged_resolutions_and_resize():
    """Resolutions that are not multiples of the 32x8 tile, and set_resolution."""
    scene = scenes.cpu_render_scene()
    tree, otree = both_trees(scene)
    cam = scenes.cpu_render_camera()
    view = S.OctreeGPUHost(tree).create_new_view(1, viewport(cam), (33, 7))
    for res in [(33, 7), (1, 1), (150, 150), (257, 129)]:
        view.set_resolution(res)
        assert view.resolution() == list(res)
        assert_frames_equal(view.render_to_host(), otree.render(oracle_camera(cam), res[0], res[1]))


def test_camera_inside_the_tree():
    scene = scenes.cpu_render_scene()
    tree, otree = both_trees(scene)
    cam = scenes.CameraSpec((20.0, 40.0, 20.0), tuple(float(v) for v in S.normalized((1.0, -0.3, 0.9))), (4.0, 4.0, 3.0), 3.0)
    view = S.OctreeGPUHost(tree).create_new_view(1, viewport(cam), (320, 240))
    assert_frames_equal(view.render_to_host(), otree.render(oracle_camera(cam), 320, 240))


def test_reload_after_edit():
    scene = scenes.cpu_render_scene()
    tree, otree = both_trees(scene)
    cam = scenes.cpu_render_camera()
    host = S.OctreeGPUHost(tree)
    view = host.create_new_view(1, viewport(cam), (150, 150))
    before = view.render_to_host()
    for t in (tree, otree):
        t.insert_at_lod((0, 48, 0), 16, 0x11FF22FF)
    view.reload()
    after = view.render_to_host()
    assert not np.array_equal(before["hit_id"], after["hit_id"])
    assert_frames_equal(after, otree.render(oracle_camera(cam), 150, 150))


def test_incremental_reload_uploads_only_written_bricks_and_matches_a_fresh_upload():
    """svx_gpu_host_reload copies the node tables plus the bricks written since the host's last upload (device brick
    index = host pool handle). After every batch of random edits the frame must equal the oracle's and the frame of a
    brand-new host (full upload) of the same tree; the pool growing past the device capacity keeps resident bricks."""
    rng = np.random.default_rng(21)
    size, dim = 128, 8
    ptree, otree = ProductOctree(size, dim), O.OracleOctree(size, dim)  # both report OctreeError as a status code
    tree = ptree.tree
    for t in (ptree, otree):
        t.insert((3, 3, 3), 0xFF0000FF)
    cam = scenes.cpu_render_camera(size)
    res = (320, 240)
    host = S.OctreeGPUHost(tree)
    view = host.create_new_view(1, viewport(cam), res)
    assert host.last_upload()["full"] and host.last_upload()["bricks"] == 1
    pool_before = host.stats()["voxel_bytes"]
    for batch in range(8):
        for _ in range(int(rng.integers(1, 60))):
            pos = tuple(int(v) for v in rng.integers(0, size, 3))
            op = int(rng.integers(0, 6))
            col = (int(rng.integers(1, 5)) * 50, int(rng.integers(0, 255)), 90, 255)
            lod = int(2 ** rng.integers(0, 3)) * dim
            status = []
            for t in (ptree, otree):
                if op <= 2:
                    status.append(t.insert(pos, col))
                elif op == 3:
                    status.append(t.insert_at_lod(pos, lod, col))
                elif op == 4:
                    status.append(t.clear(pos))
                else:
                    status.append(t.clear_at_lod(pos, lod))
            assert status[0] == status[1]
        assert tree.structure_hash() == otree.structure_hash()
        view.reload()
        up = host.last_upload()
        assert not up["full"]
        assert up["bricks"] <= 60 * 9  # an edit touches at most its own brick or the 8 a diluted brick turns into
        got = view.render_to_host()
        assert_frames_equal(got, otree.render(oracle_camera(cam), *res))
        fresh = S.OctreeGPUHost(tree).create_new_view(1, viewport(cam), res).render_to_host()
        for k in got:
            assert np.array_equal(got[k], fresh[k]), (batch, k)
    assert host.stats()["voxel_bytes"] > pool_before  # the brick pool grew (device arrays were re-allocated, contents kept)
    # one more voxel into an existing brick: exactly that brick travels
    for t in (ptree, otree):
        t.insert((3, 3, 4), 0x00FF00FF)
    view.reload()
    assert host.last_upload()["bricks"] == 1
    assert_frames_equal(view.render_to_host(), otree.render(oracle_camera(cam), *res))
    view.reload()  # nothing edited: no-op, the stats of the previous upload stay
    assert host.last_upload()["bricks"] == 1


def test_row_band_shards_compose_the_full_frame():
    """Screen sharding used for multi-GPU: the union of the shards equals the unsharded frame byte for byte."""
    scene = scenes.cpu_render_scene()
    tree = scenes.build_tree(scene, S.Octree)
    cam = scenes.cpu_render_camera()
    host = S.OctreeGPUHost(tree)
    full = host.create_new_view(1, viewport(cam), (300, 203)).render_to_host()
    for world, band in [(2, 8), (3, 16), (8, 8)]:
        acc = {k: np.zeros_like(v) for k, v in full.items()}
        rows = np.arange(203)
        for rank in range(world):
            v = host.create_new_view(1, viewport(cam), (300, 203))
            v.set_shard(rank, world, band)
            part = v.render_to_host()
            mine = (rows // band) % world == rank
            for k in acc:
                acc[k][mine] = part[k][mine]
        for k in full:
            assert np.array_equal(acc[k], full[k]), (world, band, k)


def test_render_is_deterministic_and_counts_launches():
    scene = scenes.cpu_render_scene()
    tree = scenes.build_tree(scene, S.Octree)
    view = S.OctreeGPUHost(tree).create_new_view(1, viewport(scenes.cpu_render_camera()), (640, 360))
    a = view.render_to_host()
    n0 = view.launch_count()
    b = view.render_to_host()
    assert view.launch_count() == n0 + 1
    for k in a:
        assert np.array_equal(a[k], b[k])
    f = view.render()
    assert f["kernel_ms"] > 0 and f["hit_id"]


def test_tree_loaded_from_bytes_renders_the_same_frame(tmp_path):
    """Octree::save -> Octree::load (bencode, src/octree/mod.rs:144-159): the loaded tree serialises to the same device
    buffers, so the frame is identical; the oracle's own bytes (stale pool slots included) load and render the same too."""
    scene = scenes.colonnade_scene(128, 4)
    cam = scenes.colonnade_camera(128)
    tree, otree = both_trees(scene)
    path = tmp_path / "tree.svx"
    tree.save(str(path))
    frames = []
    for t in (tree, S.Octree.load(str(path)), S.Octree.from_bytes(otree.to_bytes())):
        view = S.OctreeGPUHost(t).create_new_view(64, viewport(cam), (640, 360))
        frames.append(view.render_to_host())
    ora = otree.render(oracle_camera(cam), 640, 360)
    assert int((frames[0]["hit_id"] != S.MISS).sum()) > 10000
    for f in frames:
        assert_frames_equal(f, ora)


def test_pipelined_read_back_delivers_the_same_frames():
    """svx_view_render_to_host_async / svx_view_wait_host: frames alternate between two framebuffer slots and their
    copies overlap the next kernel; every frame must equal the synchronous render of the same pose, also when
    synchronous calls, a resize and the batch call are interleaved."""
    import torch

    scene = scenes.cpu_render_scene()
    tree = scenes.build_tree(scene, S.Octree)
    host = S.OctreeGPUHost(tree)
    cams = [scenes.cpu_render_camera(k=k) for k in range(7)]
    for res in ((320, 200), (157, 93)):
        w, h = res
        view = host.create_new_view(1, viewport(cams[0]), (64, 64))
        view.set_resolution((w, h))
        want = []
        for c in cams:
            view.set_viewport(viewport(c))
            want.append({k: v.copy() for k, v in view.render_to_host().items()})
        sets = [[torch.empty(w * h, dtype=torch.int32).pin_memory(), torch.empty(w * h, dtype=torch.int32).pin_memory(),
                 torch.empty(w * h, dtype=torch.float32).pin_memory()] for _ in range(2)]
        n0 = view.launch_count()
        got = []
        for i, c in enumerate(cams):
            view.set_viewport(viewport(c))
            view.render_to_host_async_ptr(*[b.data_ptr() for b in sets[i & 1]])
            view.wait_host(1)  # frame i-1 is complete now
            if i >= 1:
                got.append([b.numpy().copy() for b in sets[(i - 1) & 1]])
        assert view.wait_host(0) >= 0.0
        got.append([b.numpy().copy() for b in sets[(len(cams) - 1) & 1]])
        assert view.launch_count() == n0 + len(cams)
        for i, (g, wnt) in enumerate(zip(got, want)):
            assert np.array_equal(g[0].view(np.uint32).reshape(h, w), wnt["hit_id"]), (res, i)
            assert np.array_equal(g[1].view(np.uint32).reshape(h, w), wnt["albedo"]), (res, i)
            assert np.array_equal(bits(g[2].reshape(h, w)), bits(wnt["distance"])), (res, i)
        # a synchronous call right after queued frames drains them first and still sees its own pose
        view.set_viewport(viewport(cams[3]))
        view.render_to_host_async_ptr(*[b.data_ptr() for b in sets[0]])
        view.set_viewport(viewport(cams[5]))
        sync = view.render_to_host()
        assert np.array_equal(sync["hit_id"], want[5]["hit_id"])
        assert np.array_equal(sets[0][0].numpy().view(np.uint32).reshape(h, w), want[3]["hit_id"])
        # the batch call rides on the same pipeline
        batch = view.render_batch([viewport(c) for c in cams])
        for i, wnt in enumerate(want):
            assert np.array_equal(batch["hit_id"][i], wnt["hit_id"]), (res, i)
            assert np.array_equal(bits(batch["distance"][i]), bits(wnt["distance"])), (res, i)
        assert batch["kernel_ms"] > 0


def test_crawl_fast_forward_is_exact_on_a_large_tree():
    """rays that skim over a 512^3 terrain take thousands of 0.1-nudge restarts; the kernel fast-forwards
    them in closed form and must still land on the oracle's bits."""
    scene = scenes.terrain_scene(512, 8, 4321, 1, shell=4)
    tree, otree = both_trees(scene)
    cam = scenes.terrain_camera(512)
    ocam = oracle_camera(cam)
    rays = np.stack([O.pixel_ray(ocam, 320, 180, x, y) for y in range(0, 180, 3) for x in range(0, 320, 3)])
    g = S.OctreeGPUHost(tree).get_by_rays(rays)
    o = otree.get_by_rays(rays)
    assert int(o["outer_iters"].max()) > 2000  # the crawl really happens
    assert np.array_equal(g["hit"], o["hit"]) and np.array_equal(g["palette_value"], o["palette_value"])
    assert np.array_equal(bits(g["impact_point"]), bits(o["impact_point"]))
    assert np.array_equal(bits(g["distance"]), bits(o["distance"]))
    # rays from inside the tree, in every direction octant (negative steps, all binades down to the origin corner)
    rng = np.random.default_rng(5)
    origin = rng.uniform(1, 511, (3000, 3)).astype(np.float32)
    origin[:, 1] = rng.uniform(200, 511, 3000).astype(np.float32)
    d = rng.normal(size=(3000, 3)).astype(np.float32)
    d[:, 1] = np.abs(d[:, 1]) * 0.2
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    rays = np.concatenate([origin, (d / ln[:, None]).astype(np.float32)], axis=1)
    g = S.OctreeGPUHost(tree).get_by_rays(rays)
    o = otree.get_by_rays(rays)
    assert np.array_equal(g["hit"], o["hit"]) and np.array_equal(g["palette_value"], o["palette_value"])
    assert np.array_equal(bits(g["impact_point"]), bits(o["impact_point"]))


def test_random_cameras_exercise_cull_rectangle_and_prefilter():
    """Cameras at random places around (and inside, and hugging the faces of) the tree, random glass sizes and both
    glass modes: the host-side cull rectangle and the approximate root-miss prefilter must never change a pixel."""
    scene = scenes.cpu_render_scene()
    tree, otree = both_trees(scene)
    host = S.OctreeGPUHost(tree)
    rng = np.random.default_rng(2024)
    view = host.create_new_view(1, viewport(scenes.cpu_render_camera()), (96, 64))
    hits = 0
    for k in range(60):
        kind = k % 4
        if kind == 0:    # far away, cube small on screen
            origin = rng.uniform(-400, 400, 3)
        elif kind == 1:  # close to a face, cube partly off screen
            origin = rng.uniform(-20, 84, 3)
            origin[rng.integers(0, 3)] = rng.choice([-3.0, 67.0])
        elif kind == 2:  # inside the tree
            origin = rng.uniform(1, 63, 3)
        else:            # looking away / grazing
            origin = rng.uniform(-100, 164, 3)
        target = rng.uniform(-16, 80, 3) if kind != 3 else origin + rng.normal(size=3) * 50
        d = S.normalized((target - origin).astype(np.float32))
        cam = scenes.CameraSpec(tuple(float(v) for v in origin.astype(np.float32)), tuple(float(v) for v in d),
                                (float(rng.uniform(1, 8)), float(rng.uniform(1, 8)), float(rng.uniform(2, 300))),
                                float(rng.uniform(0.5, 6)), glass_at_frustum_z=bool(k % 2))
        view.set_viewport(viewport(cam))
        view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z if cam.glass_at_frustum_z else S.GLASS_AT_FOV)
        ora = otree.render(oracle_camera(cam), 96, 64)
        assert_frames_equal(view.render_to_host(), ora)
        hits += int((ora["hit_id"] != S.MISS).sum())
    assert hits > 20000


def test_clear_then_reload_matches_oracle():
    """§8(f) rank 1: Octree::clear / clear_at_lod edits followed by a re-upload (OctreeGPUView::reload)."""
    scene = scenes.cpu_render_scene()
    tree, otree = both_trees(scene)
    cam = scenes.cpu_render_camera()
    host = S.OctreeGPUHost(tree)
    view = host.create_new_view(1, viewport(cam), (320, 240))
    before = view.render_to_host()
    rng = np.random.default_rng(11)
    for t in (tree, otree):
        t.clear_at_lod((32, 32, 32), 16)       # carve a block out of the dense corner
        t.clear_at_lod((48, 48, 32), 8)
    for _ in range(300):                         # and single voxels out of the lattice and the corner
        p = tuple(int(v) for v in rng.integers(0, 64, 3))
        tree.clear(p)
        otree.clear(p)
    assert tree.structure_hash() == otree.structure_hash()
    view.reload()
    after = view.render_to_host()
    assert not np.array_equal(before["hit_id"], after["hit_id"])
    assert_frames_equal(after, otree.render(oracle_camera(cam), 320, 240))


def test_persistent_schedule_renders_the_same_frame():
    """The warp-granular persistent schedule and the static one-CTA-per-block schedule must agree byte for byte,
    including ragged resolutions, shards and repeated launches (the ticket counters ping-pong across launches)."""
    scene = scenes.cpu_render_scene()
    tree = scenes.build_tree(scene, S.Octree)
    host = S.OctreeGPUHost(tree)
    cam = scenes.cpu_render_camera()
    for res in [(640, 360), (333, 77), (35, 5)]:
        a = host.create_new_view(1, viewport(cam), res)
        b = host.create_new_view(1, viewport(cam), res)
        a.set_schedule(False)
        b.set_schedule(True)
        ref = a.render_to_host()
        for _ in range(3):
            got = b.render_to_host()
            for k in ref:
                assert np.array_equal(ref[k], got[k]), (res, k)
        a.set_shard(1, 3, 8)
        b.set_shard(1, 3, 8)
        ref, got = a.render_to_host(), b.render_to_host()
        rows = (np.arange(res[1]) // 8) % 3 == 1
        for k in ref:
            assert np.array_equal(ref[k][rows], got[k][rows]), (res, k)


def test_shaded_plane_is_the_callers_pixel():
    """The optional fourth plane = the pixel examples/cpu_render.rs:119-136 / dot_cube.rs:238-256 write: albedo scaled by
    the diffuse term of the impact normal (so the normal is checked on every hit pixel too), grey on a miss."""
    light = np.array([0.0, -1.0, 1.0], dtype=np.float32)
    light = (light / np.sqrt((light * light).sum(dtype=np.float32), dtype=np.float32)).astype(np.float32)  # V3c::normalized
    for scene, cam, res in [(scenes.cpu_render_scene(), scenes.cpu_render_camera(), (150, 150)),
                            (scenes.dot_cube_scene(128, 32), scenes.dot_cube_camera(128, zoom=True), (333, 211)),
                            (scenes.colonnade_scene(), scenes.colonnade_camera(), (320, 180))]:
        tree, otree = both_trees(scene)
        host = S.OctreeGPUHost(tree)
        view = host.create_new_view(1, viewport(cam), res)
        if cam.glass_at_frustum_z:
            view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
        ora = otree.render(oracle_camera(cam), res[0], res[1], light_normal=light)
        for persistent in (False, True):  # the shaded plane always comes from the static schedule
            view.set_schedule(persistent)
            view.set_shading(light)
            assert_frames_equal(view.render_to_host(), ora)
            shaded = view.read_shaded()
            assert np.array_equal(shaded, ora["shaded"])
            view.set_shading(None)
            assert_frames_equal(view.render_to_host(), ora)  # and the plain kernels are back
            with pytest.raises(S.OctreeError):
                view.read_shaded()
        assert (ora["shaded"] != 0xFF808080).sum() > 1000
        # with MIP maps: the LOD kernel's shaded variant
        tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
        otree.switch_albedo_mip_maps(True)
        view.reload()
        view.set_viewing_distance(80.0)
        view.set_shading(light)
        ora = otree.render(oracle_camera(cam), res[0], res[1], light_normal=light, viewing_distance=80.0)
        assert_frames_equal(view.render_to_host(), ora)
        assert np.array_equal(view.read_shaded(), ora["shaded"])
        view.set_resolution((64, 48))  # the plane follows a resize
        ora = otree.render(oracle_camera(cam), 64, 48, light_normal=light, viewing_distance=80.0)
        view.render_to_host()
        assert np.array_equal(view.read_shaded(), ora["shaded"])


@pytest.mark.parametrize("mips", [False, True], ids=["plain", "mips"])
def test_albedo_plane_is_the_colour_palette_looked_up_by_hit_id(mips):
    """svx_view_render_to_host(view, hit_id, NULL, distance) ships 8 bytes per pixel; the albedo plane it leaves out is
    svx_octree_color_palette()[hit_id & 0xFFFF] (0 for a miss or a voxel without a colour) - checked against the plane the
    three-plane call delivers, with MIP colours (which the box filter adds to the palette) too."""
    import torch

    scene = scenes.colonnade_scene()
    tree = scenes.build_tree(scene, S.Octree)
    tree.insert((3, 3, 3), None, 42)  # a voxel with user data and no colour
    if mips:
        tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
    host = S.OctreeGPUHost(tree)
    w, h = 640, 360
    view = host.create_new_view(1, viewport(scenes.colonnade_camera()), (w, h))
    if mips:
        view.set_viewing_distance(100.0)
    full = view.render_to_host()
    hit_id, dist = torch.empty(w * h, dtype=torch.int32).pin_memory(), torch.empty(w * h, dtype=torch.float32).pin_memory()
    view.render_to_host_ptr(hit_id.data_ptr(), 0, dist.data_ptr())
    ids = hit_id.numpy().view(np.uint32).reshape(h, w)
    assert np.array_equal(ids, full["hit_id"]) and np.array_equal(bits(dist.numpy().reshape(h, w)), bits(full["distance"]))
    colors = tree.color_palette()
    packed = colors[:, 0].astype(np.uint32) | (colors[:, 1].astype(np.uint32) << 8) | (colors[:, 2].astype(np.uint32) << 16) | \
        (colors[:, 3].astype(np.uint32) << 24)
    index = ids & 0xFFFF
    has_colour = (ids != 0xFFFFFFFF) & (index != 0xFFFF)
    resolved = np.where(has_colour, packed[np.minimum(index, len(packed) - 1)], 0).astype(np.uint32)
    assert has_colour.sum() > 10000 and len(np.unique(resolved)) > 3
    assert np.array_equal(resolved, full["albedo"])
