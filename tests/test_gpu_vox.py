"""`.vox` files through the whole product path on the GPU: svx_octree_load_vox (C++ loader) -> render-data upload -> viewport
kernel, against the CPU oracle fed the voxel list of the independent numpy reader. The models are the reference's own
assets/models/navigate*.vox (copied to tests/golden/vox: /root/reference does not exist on the GPU box). `-m gpu`."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
import shocovox_b200 as S
from shocovox_b200 import scenes, vox
from test_gpu_parity import assert_frames_equal, oracle_camera, viewport

pytestmark = pytest.mark.gpu
ASSETS = Path(__file__).resolve().parent / "golden" / "vox"


@pytest.mark.parametrize("name", ["navigate.vox", "navigate_x.vox", "navigate_y.vox", "navigate_z.vox"])
@pytest.mark.parametrize("brick_dim", [8, 32])
def test_vox_model_renders_like_the_oracle(name, brick_dim):
    path = ASSETS / name
    tree_size, xyz, rgba = vox.load_vox(path, brick_dimension=brick_dim)
    if tree_size < 2 * brick_dim:
        with pytest.raises(S.OctreeError):  # Octree::new refuses (the reference panics, magicavoxel.rs:273-281)
            S.Octree.load_vox_file(str(path), brick_dim)
        return
    tree = S.Octree.load_vox_file(str(path), brick_dim)
    otree = O.OracleOctree(tree_size, brick_dim)
    otree.insert_batch(xyz, rgba)
    assert tree.structure_hash() == otree.structure_hash()
    host = S.OctreeGPUHost(tree)
    hits = 0
    for k in (0, 40, 90):
        cam = scenes.cpu_render_camera(tree_size, k)
        view = host.create_new_view(64, viewport(cam), (480, 270))
        gpu = view.render_to_host()
        ora = otree.render(oracle_camera(cam), 480, 270)
        assert_frames_equal(gpu, ora)
        hits += int((ora["hit_id"] != S.MISS).sum())
    assert hits > 50


def test_vox_with_mip_strategy_renders_like_the_oracle_at_lod():
    """MIPMapStrategy::default().set_enabled(true).load_vox_file(..) (examples/minecraft.rs:57-60) and a finite viewing distance"""
    path = ASSETS / "navigate.vox"
    tree_size, xyz, rgba = vox.load_vox(path, brick_dimension=8)
    tree = S.Octree.load_vox_file(str(path), 8, mip_strategy=lambda su: su.switch_albedo_mip_maps(True))
    otree = O.OracleOctree(tree_size, 8)
    otree.switch_albedo_mip_maps(True)
    otree.insert_batch(xyz, rgba)
    assert tree.albedo_mip_map_resampling_strategy().mip_hash() == otree.mip_hash()
    cam = scenes.cpu_render_camera(tree_size)
    view = S.OctreeGPUHost(tree).create_new_view(64, viewport(cam), (480, 270))
    for vd in (3.0, 50.0):
        view.set_viewing_distance(vd)
        ora = otree.render(oracle_camera(cam), 480, 270, viewing_distance=vd)
        assert_frames_equal(view.render_to_host(), ora)
