"""MagicaVoxel import: the product's loader (csrc/vox_import.cpp behind svx_octree_load_vox; reference
src/convert/magicavoxel.rs) against the independent numpy reader (shocovox_b200/vox.py) and the oracle's tree builder. The
parsing layer is unpinned (dot_vox is not in the checkout); the placement arithmetic is pinned by the reference's rotation
KAT, and both readers must agree on synthetic scene graphs and on the reference's own assets (tests/golden/vox)."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
import shocovox_b200 as S
from product_adapter import ProductOctree
from shocovox_b200 import vox

ASSETS = Path(__file__).resolve().parent / "golden" / "vox"  # the reference's assets/models/navigate*.vox, copied unchanged


# src/convert/magicavoxel.rs:392-413
def test_matrix_parse():
    assert np.array_equal(vox.parse_rotation_matrix(4), np.eye(3, dtype=np.int64))
    m = vox.parse_rotation_matrix((1 << 0) | (2 << 2) | (0 << 4) | (1 << 5) | (1 << 6))
    assert np.array_equal(m, np.array([[0, 1, 0], [0, 0, -1], [-1, 0, 0]]))


def _model(seed, size):
    rng = np.random.default_rng(seed)
    n = 200
    pts = np.unique(rng.integers(0, size, (n, 3)), axis=0)
    idx = rng.integers(0, 255, (len(pts), 1))
    return (size, size, size), np.concatenate([pts, idx], axis=1)


def test_files_the_loaders_refuse():
    """No scene graph: the reference panics on `vox_tree.scenes[0]` (magicavoxel.rs:112). No RGBA chunk: dot_vox substitutes
    MagicaVoxel's built-in default palette, which neither reader reproduces. Both readers say so instead of inventing data."""
    size, v = _model(1, 16)
    pal = np.random.default_rng(2).integers(1, 256, (256, 4)).astype(np.uint8)
    for blob in (vox.write_vox([(size, v)], palette=pal),                                  # models only, no nTRN / nGRP / nSHP
                 vox.write_vox([(size, v)], palette=None, placements=[((0, 0, 0), None)]),  # no RGBA
                 b"VOX " + b"\x00" * 40, b"not a vox file at all"):
        with pytest.raises(ValueError):
            vox.load_vox(blob, brick_dimension=4)
        with pytest.raises(S.OctreeError) as e:
            S.Octree.load_vox_file(blob, 4)
        assert e.value.code == S.api.E_DECODE
    with pytest.raises(S.OctreeError) as e:
        S.Octree.load_vox_file("/nonexistent/file.vox", 4)
    assert e.value.code == S.api.E_IO


def test_single_model_placement():
    size, v = _model(1, 16)
    pal = np.random.default_rng(2).integers(1, 256, (256, 4)).astype(np.uint8)
    blob = vox.write_vox([(size, v)], palette=pal, placements=[((0, 0, 0), None)])
    tree_size, xyz, rgba = vox.load_vox(blob, brick_dimension=4)
    assert tree_size == 16  # translation 0 +- half size 8 -> extent 16
    # Rzup -> Lyup swaps y and z; the model is centred on the origin and shifted by the minimum corner
    want = np.stack([v[:, 0], v[:, 2], v[:, 1]], axis=1)
    assert np.array_equal(xyz.astype(np.int64), want)
    assert np.array_equal(rgba, pal[v[:, 3]])


def product_and_python_trees(blob, dim):
    """the C++ loader behind the C ABI and the numpy reader + insert_batch must build the same tree"""
    loaded = S.Octree.load_vox_file(blob, dim)
    tree_size, xyz, rgba = vox.load_vox(blob, dim)
    manual = S.Octree(tree_size, dim)
    manual.insert_batch(xyz, rgba)
    assert loaded.get_size() == tree_size
    assert loaded.structure_hash() == manual.structure_hash()
    return loaded, (tree_size, xyz, rgba)


@pytest.mark.parametrize("rot", [None, 4, (1 << 0) | (0 << 2) | (1 << 4), (1 << 0) | (2 << 2) | (1 << 5) | (1 << 6), (2 << 0) | (1 << 2) | (1 << 4)])
def test_cpp_loader_agrees_with_the_numpy_reader(rot):
    size, v = _model(11, 8)
    pal = np.random.default_rng(12).integers(1, 256, (256, 4)).astype(np.uint8)
    blob = vox.write_vox([(size, v), ((8, 8, 8), v), ((8, 8, 8), v)], palette=pal,
                         placements=[((0, 0, 0), None), ((20, -3, 7), rot), ((-11, 9, 30), rot)])
    loaded, (tree_size, xyz, rgba) = product_and_python_trees(blob, 4)
    for p, c in list(zip(xyz, rgba))[::17]:
        e = loaded.get(tuple(int(q) for q in p))
        assert e.albedo is not None and (e.albedo.r, e.albedo.g, e.albedo.b, e.albedo.a) == tuple(int(q) for q in c)


def test_scene_graph_translation_and_rotation():
    size, v = _model(3, 8)
    pal = np.random.default_rng(4).integers(1, 256, (256, 4)).astype(np.uint8)
    rot = (1 << 0) | (0 << 2) | (1 << 4)  # rows: (0,-1,0), (1,0,0), (0,0,1): a quarter turn about z
    blob = vox.write_vox([(size, v), (size, v)], palette=pal, placements=[((0, 0, 0), None), ((20, 0, 0), rot)])
    tree_size, xyz, rgba = vox.load_vox(blob, brick_dimension=4)
    assert tree_size == 32 and len(xyz) == 2 * len(v)
    assert xyz.max() < tree_size
    # the first model keeps its shape (up to the axis swap); the second is the rotated copy 20 units along x
    a = xyz[: len(v)].astype(np.int64)
    assert np.array_equal(a - a.min(axis=0), np.stack([v[:, 0], v[:, 2], v[:, 1]], axis=1) - np.stack([v[:, 0], v[:, 2], v[:, 1]], axis=1).min(axis=0))
    assert len(np.unique(xyz, axis=0)) == 2 * len(v)  # the copies do not overlap
    t = S.Octree(tree_size, 4)
    t.insert_batch(xyz, rgba)
    for p, c in list(zip(xyz, rgba))[:50]:
        e = t.get(tuple(int(q) for q in p))
        assert e.albedo is not None and (e.albedo.r, e.albedo.g, e.albedo.b, e.albedo.a) == tuple(int(q) for q in c)


@pytest.mark.parametrize("name", ["navigate.vox", "navigate_x.vox", "navigate_y.vox", "navigate_z.vox"])
def test_reference_assets_when_present(name):
    """The reference's own small models (assets/models): every voxel lands inside the tree, and the product's host
    octree and the oracle build the same tree from them."""
    path = ASSETS / name
    tree_size, xyz, rgba = vox.load_vox(path, brick_dimension=8)
    loaded = S.Octree.load_vox_file(str(path), 8)  # the product's loader (C++ behind svx_octree_load_vox)
    assert loaded.get_size() == tree_size
    assert 300 <= len(xyz) <= 100000 and tree_size in (64, 128, 256, 512)
    assert int(xyz.max()) < tree_size and int(xyz.max()) >= tree_size // 2  # the model fills more than half the extent
    a, b = O.OracleOctree(tree_size, 8), ProductOctree(tree_size, 8)
    a.insert_batch(xyz, rgba)
    b.insert_batch(xyz, rgba)
    assert a.structure_hash() == b.structure_hash() == loaded.structure_hash()
    hits = sum(1 for p in xyz[:200] if b.get(tuple(int(q) for q in p)) != (O.EMPTY,))
    assert hits == min(200, len(xyz))


def test_load_vox_file_with_a_mip_strategy_updates_mips_while_inserting():
    """MIPMapStrategy::default().set_enabled(true).load_vox_file(..) (magicavoxel.rs:207-250; examples/minecraft.rs:57-60):
    the MIPs are built incrementally by the inserts; the oracle fed the same insert sequence must end up identical."""
    size, v = _model(7, 32)
    pal = np.random.default_rng(8).integers(1, 256, (256, 4)).astype(np.uint8)
    pal[:, 3] = 255
    blob = vox.write_vox([(size, v)], palette=pal)
    blob = vox.write_vox([(size, v)], palette=pal, placements=[((0, 0, 0), None)])
    tree = S.Octree.load_vox_file(blob, 4, mip_strategy=lambda su: su.set_method_at(2, S.MIP_POINT_FILTER).set_method_at(3, S.MIP_POSTERIZE, 0.2)
                                  .set_color_similarity_thr_at(1, 0.01).switch_albedo_mip_maps(True))
    tree2 = vox.load_vox_file(blob, 4, mip_enabled=True, mip_methods={2: S.MIP_POINT_FILTER, 3: (S.MIP_POSTERIZE, 0.2)},
                              mip_color_similarity={1: 0.01})  # the numpy reader + insert_batch: same tree, same MIPs
    assert tree.structure_hash() == tree2.structure_hash()
    assert tree.albedo_mip_map_resampling_strategy().mip_hash() == tree2.albedo_mip_map_resampling_strategy().mip_hash()
    su = tree.albedo_mip_map_resampling_strategy()
    assert su.is_enabled() and su.get_method_at(2) == (S.MIP_POINT_FILTER, 0.0) and su.get_method_at(1)[0] == S.MIP_POINT_FILTER
    tree_size, xyz, rgba = vox.load_vox(blob, 4)
    o = O.OracleOctree(tree_size, 4)
    o.set_method_at(2, 1).set_method_at(3, 3, 0.2).set_color_similarity_thr_at(1, 0.01).switch_albedo_mip_maps(True)
    o.insert_batch(xyz, rgba)
    assert tree.structure_hash() == o.structure_hash() and su.mip_hash() == o.mip_hash()
    assert su.sample_root_mip(8, (0, 0, 0)).kind in (S.api.ENTRY_EMPTY, S.api.ENTRY_VISUAL)
    plain = S.Octree.load_vox_file(blob, 4)
    assert not plain.albedo_mip_map_resampling_strategy().is_enabled()
    assert plain.get_sweep((0, 0, 0), (8, 8, 8)).tobytes() == tree.get_sweep((0, 0, 0), (8, 8, 8)).tobytes()


def test_scene_graph_numbers_are_parsed_as_rust_parses_them():
    """`_t` / `_r` / `_f` go through `str::parse::<i32>()` / `<u8>()` with `unwrap()` in the reference (magicavoxel.rs:139-157):
    anything but an optional sign and digits is a panic there and a decode error here - never a guess. A file whose text
    is patched in place (same length, so the chunk sizes stay valid) must load when the number is still a number and be
    refused otherwise."""
    size, v = _model(5, 8)
    pal = np.random.default_rng(6).integers(1, 256, (256, 4)).astype(np.uint8)
    blob = vox.write_vox([(size, v), (size, v)], palette=pal, placements=[((0, 0, 0), None), ((120, -30, 700), None)])
    text = b"120 -30 700"
    assert blob.count(text) == 1
    good = S.Octree.load_vox_file(blob, 4)
    for patched, ok in [(b"+20 -30 700", True), (b"120 -30 70 ", False), (b" 20 -30 700", False), (b"120 -3x 700", False),
                        (b"120  30 700", False), (b"120 -30 7e2", False), (b"120,-30,700", False),
                        (b"1 2 3 4 500", True), (b"120 -30700 ", False), (b"120 -307000", False)]:
        data = blob.replace(text, patched)
        assert len(data) == len(blob)
        if ok:
            assert S.Octree.load_vox_file(data, 4).get_size() >= 8
        else:
            with pytest.raises(S.OctreeError) as e:
                S.Octree.load_vox_file(data, 4)
            assert e.value.code == S.api.E_DECODE
    assert good.get_size() == 1024
    # a number beyond i32 is refused as well (strtol would have saturated it)
    blob = vox.write_vox([(size, v), (size, v)], palette=pal, placements=[((0, 0, 0), None), ((99999999999, 0, 0), None)])
    with pytest.raises(S.OctreeError):
        S.Octree.load_vox_file(blob, 4)
