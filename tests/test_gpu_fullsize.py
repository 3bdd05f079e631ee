"""BASELINE.json's configs at their FULL sizes against the CPU oracle (C3 minecraft-style 1024/32 at 4K, C4
sponza-scale 2048/32 at 4K, C5 1024/8 terrain poses at 1080p). The oracle renders a bounded sample of image rows spread
evenly over the frame (sky, horizon and ground are all represented); every compared field is bit-exact. `-m gpu`."""
import numpy as np
import pytest

import bench
import oracle_lib as O
import shocovox_b200 as S
from shocovox_b200 import scenes
from test_gpu_parity import bits, oracle_camera, viewport

pytestmark = pytest.mark.gpu


def compare_rows(gpu, ora, rows, what):
    ora_rgba = ora["albedo"].view(np.uint32)[..., 0]
    assert np.array_equal(gpu["hit_id"][rows], ora["hit_id"][rows]), f"{what}: hit voxel ids differ"
    assert np.array_equal(gpu["albedo"][rows], ora_rgba[rows]), f"{what}: albedo differs"
    assert np.array_equal(bits(gpu["distance"][rows]), bits(ora["distance"][rows])), f"{what}: distance bits differ"
    assert ora["would_panic"] == 0


def render_rows(name, poses, n_rows):
    scene, cams, (w, h), _ = bench.make_workload(name)
    tree = scenes.build_tree(scene, S.Octree)
    otree = scenes.build_tree(scene, O.OracleOctree)
    assert tree.structure_hash() == otree.structure_hash()
    host = S.OctreeGPUHost(tree)
    view = host.create_new_view(64, viewport(cams[0]), (w, h))
    if cams[0].glass_at_frustum_z:
        view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
    rows = bench.sample_rows(h, n_rows)
    hits = 0
    for k in poses:
        view.set_viewport(viewport(cams[k]))
        gpu = view.render_to_host()
        ora = otree.render(oracle_camera(cams[k]), w, h, row_list=rows)
        compare_rows(gpu, ora, rows, f"{name} pose {k}")
        hits += int((ora["hit_id"][rows] != S.MISS).sum())
    return hits, len(rows) * w * len(poses)


def test_c4_sponza_4k_full_size():
    hits, rays = render_rows("sponza_4k", [0], 64)
    assert hits > rays // 4


def test_c3_minecraft_4k_full_size():
    hits, rays = render_rows("minecraft_4k", [0], 64)
    assert hits > rays // 4


def test_c5_terrain_poses_1080p_full_size():
    hits, rays = render_rows("terrain_poses_1080p", [0, 64, 128, 192], 32)
    assert hits > rays // 8
