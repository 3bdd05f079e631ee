"""The kernels' device code (shocovox_b200/csrc/traverse.cuh) compiled for the HOST and checked against the oracle - CPU only.

tests/host_mirror/host_mirror.cpp includes the very header the CUDA kernels are built from, with C++ stand-ins for its
PTX fragments, so the logic of the GPU path (transformed DDA arithmetic, mirrored brick walk, parent-index node stack,
crawl fast-forward, brick-dimension instantiations) is compared bit for bit with the restated reference
(Octree::get_by_ray, src/raytracing/raytracing_on_cpu.rs:316-565) without a GPU. Inputs are the host images of the render
data (svx_octree_render_data_nodes / _bricks, svx_render_data_ray_lut). Test infrastructure only: the product has no CPU
ray path, and what nvcc / ptxas make of the same source is the business of the `-m gpu` tests."""
import ctypes as C
import subprocess
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
import shocovox_b200 as S
from product_adapter import ProductOctree
from ray_cases import CASES
from shocovox_b200 import build as product_build
from shocovox_b200 import scenes

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "host_mirror" / "host_mirror.cpp"
HIT = np.dtype([("hit", "<u4"), ("palette_value", "<u4"), ("impact_point", "<f4", 3), ("normal", "<f4", 3), ("distance", "<f4")])


def build_mirror(name: str, defines=()):
    out = ROOT / "tests" / "host_mirror" / "build"
    out.mkdir(exist_ok=True)
    lib = out / f"lib{name}.so"
    csrc = ROOT / "shocovox_b200" / "csrc"
    deps = [SRC, csrc / "traverse.cuh", csrc / "traverse_refill.cuh", csrc / "gpu_tree.hpp", csrc / "kernels.cuh"]
    if not lib.exists() or any(d.stat().st_mtime > lib.stat().st_mtime for d in deps):
        cuda_include = Path(product_build.nvcc_path()).resolve().parent.parent / "include"
        res = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
                              *[f"-D{d}" for d in defines], f"-I{cuda_include}", f"-I{csrc}", f"-I{ROOT / 'include'}", "-o", str(lib), str(SRC)],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-3000:]
    L = C.CDLL(str(lib))
    L.svx_host_mirror_get_by_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                              C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_float]
    L.svx_host_mirror_get_by_rays.restype = C.c_int
    return L


@pytest.fixture(scope="module")
def mirror():
    """the kernel code as the library builds it (default options)"""
    return build_mirror("host_mirror")


def bits(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.where(np.isnan(a), np.uint32(0x7FC00000), a.view(np.uint32))  # NaNs compare equal (0/0 normals at a cell centre)


def mirror_rays(L, tree, rays, specialise=1, viewing_distance=None):
    """get_by_ray (viewing_distance None: the tree has no MIP maps) or get_by_ray_at_lod for every ray through the host build
    of the kernel code, on the host images of the render data"""
    rec = tree.render_data_nodes()
    mips = tree.render_data_mip_slots() if viewing_distance is not None else None
    voxels, occ = tree.render_data_bricks()
    lut = np.zeros(1024, dtype=np.uint32)
    assert S.lib().svx_render_data_ray_lut(lut.ctypes.data_as(C.c_void_p)) == 0
    rays = np.ascontiguousarray(rays, dtype=np.float32)
    out = np.zeros(len(rays), dtype=HIT)
    rc = L.svx_host_mirror_get_by_rays(rec.ctypes.data, len(rec), voxels.ctypes.data, occ.ctypes.data, len(voxels), lut.ctypes.data,
                                       tree.get_size(), tree.brick_dim(), specialise, rays.ctypes.data, len(rays), out.ctypes.data, 8,
                                       mips.ctypes.data if mips is not None else None, float(viewing_distance or 0.0))
    assert rc == 0
    return out


def assert_same_hits(got, want):
    assert np.array_equal(got["hit"] != 0, want["hit"] != 0)
    assert np.array_equal(got["palette_value"], want["palette_value"])
    assert np.array_equal(bits(got["impact_point"]), bits(want["impact_point"]))
    assert np.array_equal(bits(got["normal"]), bits(want["normal"]))
    assert np.array_equal(bits(got["distance"]), bits(want["distance"]))


def random_rays(size, n, seed):
    """the ray mix of tests/test_gpu_parity.py: outside and inside origins, a share of axis-parallel rays from lattice points"""
    rng = np.random.default_rng(seed)
    origin = rng.uniform(-1.0 * size, 2.0 * size, (n, 3)).astype(np.float32)
    inside = rng.random(n) < 0.25
    origin[inside] = rng.uniform(0, size, (int(inside.sum()), 3)).astype(np.float32)
    target = rng.uniform(0, size, (n, 3)).astype(np.float32)
    d = target - origin
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    d = (d / ln[:, None]).astype(np.float32)
    k = n // 16
    d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], k)[:, None].astype(np.float32)
    origin[:k] = np.round(origin[:k])
    return np.concatenate([origin, d], axis=1)


SCENES = {
    "cpu_render_64_8": lambda: scenes.cpu_render_scene(64, 8),          # brick-8 instantiation
    "cpu_render_32_1": lambda: scenes.cpu_render_scene(32, 1),          # generic code, one-voxel bricks, depth 5
    "cpu_render_32_2": lambda: scenes.cpu_render_scene(32, 2),
    "cpu_render_64_4": lambda: scenes.cpu_render_scene(64, 4),
    "dot_cube_128_32": lambda: scenes.dot_cube_scene(128, 32),          # brick-32 instantiation
    "criterion_128_8": lambda: scenes.criterion_scene(128, 8, 48),
    "colonnade_256_8": scenes.colonnade_scene,                          # UniformLeaf nodes and Solid bricks next to Parted ones
    "terrain_blocky_128_8": lambda: scenes.terrain_scene(128, 8, 1234, 4),
    "terrain_256_8_shell": lambda: scenes.terrain_scene(256, 8, 4321, 1, shell=3),  # long crawls over empty space
}


@pytest.mark.parametrize("name", list(SCENES))
def test_random_rays_bit_exact_on_the_host_build_of_the_kernel_code(mirror, name):
    scene = SCENES[name]()
    tree, otree = scenes.build_tree(scene, S.Octree), scenes.build_tree(scene, O.OracleOctree)
    assert tree.structure_hash() == otree.structure_hash()
    rays = random_rays(scene.tree_size, 20000, zlib.crc32(name.encode()) % 1000)
    want = otree.get_by_rays(rays)
    assert int(want["would_panic"].sum()) == 0 and want["hit"].sum() > 300
    assert_same_hits(mirror_rays(mirror, tree, rays), want)
    if tree.brick_dim() in (8, 32):  # the generic code on the same tree: both instantiations are what the library launches
        assert_same_hits(mirror_rays(mirror, tree, rays[:5000], specialise=0), want[:5000])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_edge_case_rays(mirror, case):
    """The 17 deterministic rays of src/raytracing/tests.rs:253-813 (deep_stack exercises the ring stack's overflow through
    the parent index, cube_flaps must miss, detailed_brick_z_edge_error pins a normal)."""
    p, o = ProductOctree(case["size"], case["dim"]), O.OracleOctree(case["size"], case["dim"])
    case["build"](p)
    case["build"](o)
    ray = np.concatenate([np.asarray(case["origin"], np.float32), np.asarray(case["direction"], np.float32)])[None]
    assert_same_hits(mirror_rays(mirror, p.tree, ray), o.get_by_rays(ray))


def test_edits_are_followed(mirror):
    """inserts, insert_at_lod and clears between queries: the render-data images follow the tree"""
    rng = np.random.default_rng(5)
    ptree, otree = ProductOctree(64, 8), O.OracleOctree(64, 8)  # the adapter gives both the same insert / clear signature
    rays = random_rays(64, 4000, 77)
    for round_ in range(4):
        for _ in range(300):
            p = tuple(int(v) for v in rng.integers(0, 64, 3))
            c = (int(rng.integers(1, 255)) << 24) | 0x50A0FF  # 0xRRGGBBAA
            op = int(rng.integers(0, 10))
            for t in (ptree, otree):
                if op < 7:
                    t.insert(p, c)
                elif op < 8:
                    t.insert_at_lod(tuple((v // 8) * 8 for v in p), 8, c)
                else:
                    t.clear(p)
        assert_same_hits(mirror_rays(mirror, ptree.tree, rays), otree.get_by_rays(rays))


F32_MAX = 3.4028234663852886e38
LOD_SCENES = {
    "cpu_render_64_8": lambda: scenes.cpu_render_scene(64, 8),
    "cpu_render_32_1": lambda: scenes.cpu_render_scene(32, 1),
    "dot_cube_128_32": lambda: scenes.dot_cube_scene(128, 32),
    "colonnade_256_8": scenes.colonnade_scene,
    "terrain_256_8_shell": lambda: scenes.terrain_scene(256, 8, 4321, 1, shell=3),  # deeper than the 4-entry ring stack
}


@pytest.mark.parametrize("name", list(LOD_SCENES))
def test_level_of_detail_branch_bit_exact(mirror, name):
    """get_by_ray_at_lod (raytracing_on_cpu.rs:325-386: the node's MIP brick answers when the node is far enough away) through
    traverse<LOD = true>: the drifting mip_level, the bracketing pre-test of the LOD condition, the crawl fast-forward under
    LOD - for viewing distances on both sides of every decision, and the degenerate ones."""
    scene = LOD_SCENES[name]()
    tree, otree = scenes.build_tree(scene, S.Octree), scenes.build_tree(scene, O.OracleOctree)
    tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
    otree.switch_albedo_mip_maps(True)
    assert tree.albedo_mip_map_resampling_strategy().mip_hash() == otree.mip_hash()
    rays = random_rays(scene.tree_size, 6000, 3 + zlib.crc32(name.encode()) % 1000)
    probes = 0
    for vd in (F32_MAX, 1000.0, 50.0, 3.0, 0.5, 0.0, -1.0, float("inf"), float("nan")):
        want = otree.get_by_rays_at_lod(rays, vd)
        assert_same_hits(mirror_rays(mirror, tree, rays, viewing_distance=vd), want)
        probes += int(want["hit"].sum())
    assert probes > 1000
    assert_same_hits(mirror_rays(mirror, tree, rays[:2000], specialise=0, viewing_distance=50.0), otree.get_by_rays_at_lod(rays[:2000], 50.0))


# ---- the resumable form of the traversal (traverse_refill.cuh), what the lane-refill schedule runs ------------------------

def test_traverse_refill_header_is_what_the_generator_makes_of_traverse_cuh():
    """traverse_refill.cuh is derived text: tools/make_traverse_refill.py cuts traverse() out of traverse.cuh and turns its
    locals into a state record. A change of the traversal that was not regenerated fails here."""
    sys.path.insert(0, str(ROOT / "tools"))
    import make_traverse_refill
    committed = (ROOT / "shocovox_b200" / "csrc" / "traverse_refill.cuh").read_text()
    assert make_traverse_refill.generate() == committed, "run python tools/make_traverse_refill.py"


@pytest.mark.parametrize("quantum", [1, 3, 24])
def test_suspended_and_resumed_traversal_is_bit_exact(quantum):
    """traverse_begin + traverse_resumable called until it stops reporting WALK_SUSPENDED, with `quantum` node-loop iterations
    per call (1 suspends at every possible point): every field equals the oracle's, plain and at LOD."""
    L = build_mirror(f"host_mirror_resumable_{quantum}", defines=(f"SVX_MIRROR_RESUMABLE={quantum}",))
    for name in ("cpu_render_64_8", "cpu_render_32_1", "dot_cube_128_32", "colonnade_256_8", "terrain_256_8_shell"):
        scene = SCENES[name]()
        tree, otree = scenes.build_tree(scene, S.Octree), scenes.build_tree(scene, O.OracleOctree)
        rays = random_rays(scene.tree_size, 6000, 11 + zlib.crc32(name.encode()) % 1000)
        want = otree.get_by_rays(rays)
        assert want["hit"].sum() > 100
        assert_same_hits(mirror_rays(L, tree, rays), want)
        assert_same_hits(mirror_rays(L, tree, rays[:2000], specialise=0), want[:2000])
        tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
        otree.switch_albedo_mip_maps(True)
        for vd in (F32_MAX, 50.0, 3.0, 0.0):
            assert_same_hits(mirror_rays(L, tree, rays[:3000], viewing_distance=vd), otree.get_by_rays_at_lod(rays[:3000], vd))
