"""The product's host octree must build the SAME tree shape as the oracle from the same insert sequence:
the shape decides which f32 path a ray takes. Compared with a key-order independent structural hash (node kinds,
occupancy bits, brick kinds and contents, palettes) and with get() sweeps."""
import numpy as np
import pytest

import oracle_lib as O
import shocovox_b200 as S
from product_adapter import ProductOctree
from shocovox_b200 import scenes
from ray_cases import CASES


def _same(a, b, size):
    assert a.structure_hash() == b.structure_hash()
    n = min(size, 32)
    sa, sb = a.get_sweep((0, 0, 0), (n, n, n)), b.get_sweep((0, 0, 0), (n, n, n))
    assert np.array_equal(sa["kind"], sb["kind"])
    assert np.array_equal(sa["rgba"], sb["rgba"])
    assert np.array_equal(sa["data"], sb["data"])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_edge_case_trees_have_identical_shape(case):
    a, b = O.OracleOctree(case["size"], case["dim"]), ProductOctree(case["size"], case["dim"])
    case["build"](a)
    case["build"](b)
    _same(a, b, case["size"])


@pytest.mark.parametrize("scene_fn", [scenes.cpu_render_scene, lambda: scenes.cpu_render_scene(128, 8),
                                      lambda: scenes.cpu_render_scene(32, 1), lambda: scenes.cpu_render_scene(32, 2),
                                      lambda: scenes.dot_cube_scene(64, 8), lambda: scenes.dot_cube_scene(128, 32),
                                      lambda: scenes.criterion_scene(128, 8, 48), scenes.colonnade_scene,
                                      lambda: scenes.terrain_scene(64, 8, 1234, 4), lambda: scenes.terrain_scene(64, 4, 4321, 1)],
                         ids=["cpu_render_64_8", "cpu_render_128_8", "cpu_render_32_1", "cpu_render_32_2", "dot_cube_64_8",
                              "dot_cube_128_32", "criterion_128_8", "colonnade_256_8", "terrain_blocky_64_8", "terrain_64_4"])
def test_scene_trees_have_identical_shape(scene_fn):
    sc = scene_fn()
    a = scenes.build_tree(sc, O.OracleOctree)
    b = scenes.build_tree(sc, ProductOctree)
    _same(a, b, sc.tree_size)


@pytest.mark.parametrize("size,dim,seed", [(8, 1, 0), (8, 2, 1), (16, 2, 2), (16, 4, 3), (32, 4, 4), (32, 8, 5), (64, 8, 6)])
def test_random_edit_sequences_have_identical_shape(size, dim, seed):
    """Random interleaving of insert / insert_at_lod / update / clear / clear_at_lod with few colours, so that
    simplification, uniform-leaf splitting, whole-node overwrites and node removal all trigger (seeded)."""
    rng = np.random.default_rng(seed)
    a, b = O.OracleOctree(size, dim), ProductOctree(size, dim)
    colors = [0xFF0000FF, 0x00FF00FF, 0x0000FFFF]
    for step in range(1500):
        p = tuple(int(v) for v in rng.integers(0, size, 3))
        op = int(rng.integers(0, 10))
        c = colors[int(rng.integers(0, len(colors)))]
        if op < 6:
            ra, rb = a.insert(p, c), b.insert(p, c)
        elif op < 7:
            lod = int(2 ** rng.integers(1, 4))
            q = tuple((v // lod) * lod for v in p) if rng.integers(0, 2) else p
            ra, rb = a.insert_at_lod(q, lod, c), b.insert_at_lod(q, lod, c)
        elif op < 8:
            if rng.integers(0, 3) == 0:
                lod = int(2 ** rng.integers(1, 4))
                q = tuple((v // lod) * lod for v in p) if rng.integers(0, 2) else p
                ra, rb = a.clear_at_lod(q, lod), b.clear_at_lod(q, lod)
            else:
                ra, rb = a.clear(p), b.clear(p)
        elif op < 9:
            d = int(rng.integers(1, 4))
            ra, rb = a.update(p, None, d), b.update(p, None, d)
        else:
            d = int(rng.integers(1, 3))
            ra, rb = a.insert(p, c, d), b.insert(p, c, d)
        assert ra == rb
        if step % 250 == 0:
            assert a.structure_hash() == b.structure_hash(), step
    _same(a, b, size)


def test_dense_fill_collapses_identically():
    for size, dim in [(8, 2), (16, 4), (16, 1)]:
        a, b = O.OracleOctree(size, dim), ProductOctree(size, dim)
        for t in (a, b):
            for x in range(size):
                for y in range(size):
                    for z in range(size):
                        t.insert((x, y, z), 0x808080FF)
        _same(a, b, size)
        assert a.node_count() == b.node_count()
