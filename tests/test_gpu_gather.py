"""The tile-sharded gather (csrc/multi_gpu.cu) against the single-GPU frame, byte for byte.

The protocol - peers store into the root's framebuffer, go / done flags on the device - does not care whether the members
sit on different GPUs, so most of it is exercised here on ONE device (several views, each with its own stream and its own
replica of the tree): that is what a 1-GPU test box can witness. The cross-device variants are in test_gpu_multi.py. `-m gpu`."""
import os

import numpy as np
import pytest

import shocovox_b200 as S
from shocovox_b200 import api, scenes
from test_gpu_parity import bits, viewport

pytestmark = pytest.mark.gpu

PLANES = ("hit_id", "albedo", "distance")


def assert_same_frame(got, want, what=""):
    for k in PLANES:
        assert np.array_equal(bits(got[k]) if k == "distance" else got[k], bits(want[k]) if k == "distance" else want[k]), (what, k)


@pytest.fixture(scope="module")
def tree():
    return scenes.build_tree(scenes.cpu_render_scene(), S.Octree)


def whole_frames(tree, cams, res):
    host = S.OctreeGPUHost(tree)
    view = host.create_new_view(1, viewport(cams[0]), res)
    out = []
    for c in cams:
        view.set_viewport(viewport(c))
        out.append(view.render_to_host())
    return out


@pytest.mark.parametrize("wire", [S.WIRE_THREE_PLANES, S.WIRE_ID_DISTANCE], ids=["12B", "8B"])
@pytest.mark.parametrize("world,band,res", [(2, 8, (640, 360)), (4, 8, (300, 203)), (3, 16, (317, 117)), (8, 4, (256, 256))])
def test_gather_members_on_one_device_assemble_the_single_gpu_frame(tree, wire, world, band, res):
    cams = [scenes.cpu_render_camera(k=7 * i) for i in range(4)]
    want = whole_frames(tree, cams, res)
    hosts = [S.OctreeGPUHost(tree) for _ in range(world)]  # one replica per member, like one per GPU
    views = [h.create_new_view(1, viewport(cams[0]), res) for h in hosts]
    views[0].gather_open(world, band, wire, export=False)
    for r in range(1, world):
        views[r].gather_join_local(r, views[0])
    assert views[0].gather_info()["role"] == "root" and views[1].gather_info() == {"role": "peer", "rank": 1, "world": world, "frames": 0}
    for i, c in enumerate(cams):
        # any submission order works: peers wait on the device for the root's go, the root for the peers' done
        order = list(range(world)) if i % 2 else list(range(world - 1, -1, -1))
        for r in order:
            views[r].set_viewport(viewport(c))
            views[r].render(sync=False)
        got = views[0].read_frame()
        assert_same_frame(got, want[i], f"frame {i}")
    assert views[0].gather_info()["frames"] == len(cams)
    for r in range(world - 1, -1, -1):
        views[r].gather_close()
    # after the gather every view renders whole frames into its own framebuffer again
    assert_same_frame(views[1].render_to_host(), want[-1], "after close")


def test_gather_refuses_what_would_break_it(tree):
    cam = scenes.cpu_render_camera()
    host = S.OctreeGPUHost(tree)
    root, peer, other = (host.create_new_view(1, viewport(cam), r) for r in ((64, 48), (64, 48), (32, 48)))
    with pytest.raises(S.OctreeError):
        peer.gather_join_local(1, root)  # the root is not open
    root.gather_open(2, 8, export=False)
    for bad in (lambda: other.gather_join_local(1, root),     # another resolution
                lambda: peer.gather_join_local(0, root),      # rank 0 is the root
                lambda: peer.gather_join_local(2, root),      # beyond the world
                lambda: root.set_resolution((32, 32)),
                lambda: root.set_shard(1, 2, 8),
                lambda: root.set_compact_rows(True),
                lambda: root.set_shading((0.0, -1.0, 0.0)),
                lambda: root.gather_open(4, 8, export=False)):  # another shape while open
        with pytest.raises(S.OctreeError) as e:
            bad()
        assert e.value.code == api.E_INVALID_ARGUMENT
    peer.gather_join_local(1, root)
    with pytest.raises(S.OctreeError):
        peer.render_to_host_async_ptr(0, 0, 0)
    peer.gather_close()
    root.gather_close()


def test_a_missing_peer_is_a_timeout_not_a_hang(tree):
    os.environ["SVX_GATHER_TIMEOUT_MS"] = "200"
    try:
        host = S.OctreeGPUHost(tree)
        root = host.create_new_view(1, viewport(scenes.cpu_render_camera()), (64, 48))
    finally:
        del os.environ["SVX_GATHER_TIMEOUT_MS"]
    root.gather_open(2, 8, export=False)
    with pytest.raises(S.OctreeError) as e:
        root.render(sync=True)  # nobody renders rank 1's rows
    assert e.value.code == api.E_TIMEOUT
    root.gather_close()
    assert (root.render_to_host()["hit_id"] != S.MISS).any()  # the view is usable again


@pytest.mark.parametrize("wire", [S.WIRE_THREE_PLANES, S.WIRE_ID_DISTANCE], ids=["12B", "8B"])
def test_svx_multi_on_one_device(tree, wire):
    cams = [scenes.cpu_render_camera(k=3 * i) for i in range(5)]
    res = (320, 203)
    want = whole_frames(tree, cams, res)
    m = S.MultiGPU(tree, [0, 0, 0], viewport(cams[0]), res, rows_per_band=8, wire=wire)
    for i, c in enumerate(cams[:3]):
        m.set_viewport(viewport(c))
        assert_same_frame(m.read_root_frame(), want[i], f"gathered frame {i}")
        assert_same_frame(m.render_to_host(), want[i], f"host-assembled frame {i}")
    batch = m.render_poses([viewport(c) for c in cams])
    for i in range(len(cams)):
        assert_same_frame({k: batch[k][i] for k in PLANES}, want[i], f"pose {i}")


def test_sharded_views_assemble_one_frame_in_shared_host_planes(tree):
    """svx_view_render_to_host on a local shard copies exactly the rows it owns: several GPUs (here: views) fill one set of host planes."""
    cam = scenes.cpu_render_camera()
    res = (300, 203)  # the last band is partial
    want = whole_frames(tree, [cam], res)[0]
    host = S.OctreeGPUHost(tree)
    for world, band in [(2, 8), (3, 16), (8, 8)]:
        planes = {"hit_id": np.full((res[1], res[0]), 0xDEADBEEF, np.uint32), "albedo": np.full((res[1], res[0]), 0xDEADBEEF, np.uint32),
                  "distance": np.full((res[1], res[0]), -1.0, np.float32)}
        for rank in range(world):
            v = host.create_new_view(1, viewport(cam), res)
            v.set_shard(rank, world, band)
            v.render_to_host(planes["hit_id"], planes["albedo"], planes["distance"])
        assert_same_frame(planes, want, f"world {world} band {band}")


def test_degenerate_cameras_and_rays_are_rejected_not_rendered(tree):
    """a direction parallel to `up` makes `up x direction` zero and every ray NaN; NaN / zero ray directions
    never step in the DDA. The reference debug-asserts (spatial/raytracing/mod.rs:14-16); here: SVX_E_INVALID_ARGUMENT."""
    host = S.OctreeGPUHost(tree)
    good = viewport(scenes.cpu_render_camera())
    for direction in [(0.0, 1.0, 0.0), (0.0, -1.0, 0.0), (0.0, 0.0, 0.0), (float("nan"), 0.0, 1.0), (float("inf"), 0.0, 0.0)]:
        with pytest.raises(S.OctreeError) as e:
            host.create_new_view(1, S.Viewport(good.origin, direction, good.frustum, good.fov), (32, 32))
        assert e.value.code == api.E_INVALID_ARGUMENT
    view = host.create_new_view(1, good, (32, 32))
    with pytest.raises(S.OctreeError):
        view.set_viewport(S.Viewport((float("nan"), 0.0, 0.0), good.direction, good.frustum, good.fov))
    for ray in [(1, 1, 1, 0, 0, 0), (1, 1, 1, float("nan"), 0, 1), (float("inf"), 1, 1, 0, 0, 1)]:
        with pytest.raises(S.OctreeError) as e:
            host.get_by_rays(np.array([ray], np.float32))
        assert e.value.code == api.E_INVALID_ARGUMENT
    # extreme but finite inputs terminate (non-finite entry points are misses)
    rays = np.array([[3e38, 3e38, 3e38, -0.57735026, -0.57735026, -0.57735026], [-3e38, 5.0, 5.0, 1.0, 0.0, 0.0],
                     [1e30, 1e30, -1e30, -0.57735026, -0.57735026, 0.57735026]], np.float32)
    out = host.get_by_rays(rays)
    assert out.shape[0] == 3


def test_heaviest_first_block_order_renders_the_same_frames(tree):
    """SVX_CTA_ORDER=2 forces the recorded block order (kernels.cuh: FrameParams::cta_order) on every static launch: frame 1
    records the cost of every 16x8 block, every later frame is dispatched heaviest-first by the previous frame's costs - also
    when the pose has changed in between. Whole frames, shards and gather members: bytes identical to the raster order."""
    cams = [scenes.cpu_render_camera(k=11 * i) for i in range(5)]
    res = (640, 363)
    want = whole_frames(tree, cams, res)
    os.environ["SVX_CTA_ORDER"] = "2"
    try:
        host = S.OctreeGPUHost(tree)
        whole = host.create_new_view(1, viewport(cams[0]), res)
        members = [S.OctreeGPUHost(tree).create_new_view(1, viewport(cams[0]), res) for _ in range(4)]
        shards = [host.create_new_view(1, viewport(cams[0]), res) for _ in range(4)]
    finally:
        del os.environ["SVX_CTA_ORDER"]
    for r, v in enumerate(shards):
        v.set_shard(r, 4, 8)
    members[0].gather_open(4, 8, S.WIRE_ID_DISTANCE, export=False)
    for r in range(1, 4):
        members[r].gather_join_local(r, members[0])
    for i, c in enumerate(cams):
        whole.set_viewport(viewport(c))
        assert_same_frame(whole.render_to_host(), want[i], f"whole frame {i}")
        planes = {"hit_id": np.zeros((res[1], res[0]), np.uint32), "albedo": np.zeros((res[1], res[0]), np.uint32),
                  "distance": np.zeros((res[1], res[0]), np.float32)}
        for v in shards:
            v.set_viewport(viewport(c))
            v.render_to_host(planes["hit_id"], planes["albedo"], planes["distance"])
        assert_same_frame(planes, want[i], f"shards {i}")
        for v in reversed(members):
            v.set_viewport(viewport(c))
            v.render(sync=False)
        assert_same_frame(members[0].read_frame(), want[i], f"gather {i}")
    for v in reversed(members):
        v.gather_close()


def test_handles_may_be_freed_in_any_order():
    """A view keeps its host alive and a host its octree (capi_internal.hpp): finalisers that run in no particular order -
    Python's cycle collector after an exception, for one - must not leave dangling handles. Raw C-ABI calls."""
    import ctypes as C

    L = S.lib()
    scene = scenes.cpu_render_scene()
    cam = scenes.cpu_render_camera()
    want = whole_frames(scenes.build_tree(scene, S.Octree), [cam], (96, 64))[0]
    tree = C.c_void_p()
    assert L.svx_octree_new(scene.tree_size, scene.brick_dim, C.byref(tree)) == 0
    xyz, rgba = np.ascontiguousarray(scene.xyz, np.uint32), np.ascontiguousarray(scene.rgba, np.uint8)
    assert L.svx_octree_insert_batch(tree, xyz.ctypes.data, rgba.ctypes.data, None, len(xyz)) == 0
    host, view = C.c_void_p(), C.c_void_p()
    assert L.svx_gpu_host_create(tree, 0, C.byref(host)) == 0
    vp = viewport(cam)._c()
    assert L.svx_gpu_host_create_view(host, 1, C.byref(vp), 96, 64, C.byref(view)) == 0
    L.svx_octree_free(tree)     # the octree first,
    L.svx_gpu_host_free(host)   # then the host: the view still renders
    got = {"hit_id": np.empty((64, 96), np.uint32), "albedo": np.empty((64, 96), np.uint32), "distance": np.empty((64, 96), np.float32)}
    assert L.svx_view_render_to_host(view, got["hit_id"].ctypes.data, got["albedo"].ctypes.data, got["distance"].ctypes.data) == 0
    assert_same_frame(got, want)
    L.svx_view_free(view)       # the last reference takes host and octree with it


def test_a_root_that_goes_first_takes_its_in_process_peers_out_of_the_gather(tree):
    """svx_view_gather_join_local peers store through the root view's own device pointers. Closing - or freeing - the root
    before them detaches them (they render whole frames into their own framebuffers again) instead of leaving them
    pointing into a freed frame. Raw C-ABI calls, so that no Python object keeps the root alive."""
    import ctypes as C

    L = S.lib()
    cam = scenes.cpu_render_camera()
    res = (200, 120)
    want = whole_frames(tree, [cam], res)[0]
    vp = viewport(cam)._c()
    for how in ("close", "free"):
        hosts, views = [C.c_void_p() for _ in range(3)], [C.c_void_p() for _ in range(3)]
        for h, v in zip(hosts, views):
            assert L.svx_gpu_host_create(tree._h, 0, C.byref(h)) == 0
            assert L.svx_gpu_host_create_view(h, 1, C.byref(vp), res[0], res[1], C.byref(v)) == 0
        assert L.svx_view_gather_open(views[0], 3, 8, S.WIRE_THREE_PLANES, None) == 0
        for r in (1, 2):
            assert L.svx_view_gather_join_local(views[r], r, views[0]) == 0
        for r in (2, 1, 0):
            assert L.svx_view_render(views[r], None) == 0
        if how == "close":
            assert L.svx_view_gather_close(views[0]) == 0
        else:
            L.svx_view_free(views[0])
        role, rank, world, frames = C.c_int32(-1), C.c_uint32(9), C.c_uint32(9), C.c_uint32(9)
        for r in (1, 2):
            assert L.svx_view_gather_info(views[r], C.byref(role), C.byref(rank), C.byref(world), C.byref(frames)) == 0
            assert (role.value, rank.value, world.value) == (0, 0, 1)
            got = {"hit_id": np.empty(res[::-1], np.uint32), "albedo": np.empty(res[::-1], np.uint32), "distance": np.empty(res[::-1], np.float32)}
            assert L.svx_view_render_to_host(views[r], got["hit_id"].ctypes.data, got["albedo"].ctypes.data, got["distance"].ctypes.data) == 0
            assert_same_frame(got, want, f"{how}: former peer {r}")
            assert L.svx_view_gather_close(views[r]) == 0  # closing twice is fine
        for r in ((0, 1, 2) if how == "close" else (1, 2)):
            L.svx_view_free(views[r])
        for h in hosts:
            L.svx_gpu_host_free(h)
