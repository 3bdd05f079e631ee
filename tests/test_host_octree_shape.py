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


def _blocky_edits(rng, size, n):
    """Edits that keep bricks close to 'every aligned 2x2x2 block is one value' - the state simplify() tests for after
    every insert (update/mod.rs:884-980) and both implementations answer from a per-brick block cache: blocks completed voxel
    by voxel in random order, blocks broken and repaired, boxes written with insert_at_lod, boxes cleared."""
    ops = []
    colors = [0xC03020FF, 0x20C030FF, 0x3020C0FF]
    for _ in range(n):
        bx, by, bz = (int(v) * 2 for v in rng.integers(0, size // 2, 3))
        c = colors[int(rng.integers(0, len(colors)))]
        k = int(rng.integers(0, 10))
        if k < 5:      # a whole block, voxel by voxel, in random order
            for o in rng.permutation(8):
                ops.append(("insert", (bx + (int(o) & 1), by + ((int(o) >> 1) & 1), bz + (int(o) >> 2)), c))
        elif k < 7:    # break a block: one voxel of another colour / one voxel cleared
            p = (bx + int(rng.integers(0, 2)), by + int(rng.integers(0, 2)), bz + int(rng.integers(0, 2)))
            ops.append(("insert", p, colors[(colors.index(c) + 1) % 3]) if k == 5 else ("clear", p))
        elif k < 9:    # an aligned or unaligned box
            s = int(rng.choice([2, 3, 4, 8]))
            p = (bx, by, bz) if k == 7 else (bx + 1, by, bz + 1)
            ops.append(("insert_at_lod", tuple(min(v, size - 1) for v in p), s, c))
        else:
            ops.append(("clear_at_lod", (bx, by, bz), int(rng.choice([2, 4]))))
    return ops


def _apply(t, op):
    if op[0] == "insert":
        t.insert(op[1], op[2])
    elif op[0] == "clear":
        t.clear(op[1])
    elif op[0] == "insert_at_lod":
        t.insert_at_lod(op[1], op[2], op[3])
    else:
        t.clear_at_lod(op[1], op[2])


# structure hashes of the trees these edits built BEFORE the block caches existed (both implementations scanning the brick on
# every question, the way the reference does): the caches must not change a single tree
FROZEN_BLOCKY = {(16, 4, 1): 0xE20DAA9B23E29AB9, (32, 8, 2): 0x49277D8BC990C71C, (64, 32, 3): 0x362E8EF8FD25F175,
                 (32, 16, 4): 0x650A39239517F6E0, (8, 2, 5): 0x56A71F7BB99BA093}


@pytest.mark.parametrize("size,dim,seed", list(FROZEN_BLOCKY))
def test_blocky_edit_sequences_build_the_trees_they_always_built(size, dim, seed):
    rng = np.random.default_rng(seed)
    ops = _blocky_edits(rng, size, 260)
    a, b = O.OracleOctree(size, dim), ProductOctree(size, dim)
    for i, op in enumerate(ops):
        _apply(a, op)
        _apply(b, op)
        if i % 97 == 0:
            assert a.structure_hash() == b.structure_hash(), (i, op)
    _same(a, b, size)
    assert a.structure_hash() == FROZEN_BLOCKY[(size, dim, seed)]


def test_blocky_terrain_builds_the_tree_it_always_built():
    """the 32^3-brick blocky terrain of BASELINE config 3 at 128^3, voxel by voxel: frozen like the sequences above"""
    sc = scenes.terrain_scene(128, 32, 1234, 4, shell=8, name="minecraft")
    a, b = scenes.build_tree(sc, O.OracleOctree), scenes.build_tree(sc, ProductOctree)
    assert a.structure_hash() == b.structure_hash() == 0xB90D1F317CDCEC69
