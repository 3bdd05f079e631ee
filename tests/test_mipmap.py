"""MIP maps (src/octree/mipmap.rs): the reference's known-answer tests (src/octree/tests.rs:154-567, `mod mipmap_tests`)
restated against the CPU oracle AND the product's host octree, plus oracle == product digests of the MIP bricks under
random edit sequences and for every resampling method. Paths relative to /root/reference/."""
import math

import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OracleOctree, entry_key as K
from product_adapter import ProductOctree

RED, GREEN, BLUE = 0xFF0000FF, 0x00FF00FF, 0x0000FFFF
OOB_OCTANT = 8
BOX, POINT, POINT_BD, POSTERIZE, POSTERIZE_BD = 0, 1, 2, 3, 4  # MIPResamplingMethods, types.rs:106-139


@pytest.fixture(params=["oracle", "product"])
def Tree(request):
    return OracleOctree if request.param == "oracle" else ProductOctree


def f32(v):
    return float(np.float32(v))


def gamma_mix(n):
    """`((255_f32.powf(2.) / n).sqrt() as u32)` of the reference tests (tests.rs:163-167)"""
    return int(np.sqrt(np.float32(255.0) * np.float32(255.0) / np.float32(n), dtype=np.float32))


def albedo_of(entry_key):
    assert entry_key[0] in (O.VISUAL, O.COMPLEX), entry_key
    return tuple(entry_key[1])


def fill_rg(t, base=(0, 0, 0)):
    bx, by, bz = base
    for (x, y, z), c in [((0, 0, 0), RED), ((0, 0, 1), GREEN), ((0, 1, 0), RED), ((0, 1, 1), GREEN), ((1, 0, 0), RED),
                         ((1, 0, 1), GREEN)]:
        assert t.insert((bx + x, by + y, bz + z), c) == O.OK


# tests.rs:160-204
def test_mixed_mip_lvl1(Tree):
    m = gamma_mix(2)
    t = Tree(2, 1)
    t.set_auto_simplify(False)
    t.switch_albedo_mip_maps(True).set_method_at(1, BOX)
    fill_rg(t)
    assert albedo_of(t.sample_root_mip(OOB_OCTANT, (0, 0, 0))) == (m, m, 0, 255)
    assert m == 180


# tests.rs:207-282
def test_simple_solid_mip_lvl2_where_dim_is_2(Tree):
    t = Tree(4, 2)
    t.set_auto_simplify(False)
    t.switch_albedo_mip_maps(True).set_method_at(1, BOX)
    for p in [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1)]:
        assert t.insert(p, RED) == O.OK
    assert albedo_of(t.sample_root_mip(OOB_OCTANT, (0, 0, 0))) == (255, 0, 0, 255)
    for p in [(0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]:
        assert t.sample_root_mip(OOB_OCTANT, p) == (O.EMPTY,)


# tests.rs:285-362
def test_mixed_mip_lvl2_where_dim_is_2(Tree):
    m = gamma_mix(2)
    t = Tree(4, 2)
    t.set_auto_simplify(False)
    t.switch_albedo_mip_maps(True).set_method_at(1, BOX)
    fill_rg(t)
    assert albedo_of(t.sample_root_mip(OOB_OCTANT, (0, 0, 0))) == (m, m, 0, 255)
    for p in [(0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]:
        assert t.sample_root_mip(OOB_OCTANT, p) == (O.EMPTY,)


def _dim4_scene(t):
    fill_rg(t)
    for p, c in [((8, 0, 0), RED), ((8, 0, 1), GREEN), ((8, 1, 0), BLUE), ((8, 1, 1), GREEN), ((9, 1, 0), RED),
                 ((9, 0, 1), BLUE)]:
        assert t.insert(p, c) == O.OK


def _dim4_asserts(t):
    m2, m3 = gamma_mix(2), gamma_mix(3)
    rg_mix, rgb_mix = (m2, m2, 0, 255), (m3, m3, m3, 255)
    assert albedo_of(t.sample_root_mip(0, (0, 0, 0))) == rg_mix      # child at 0,0,0
    assert albedo_of(t.sample_root_mip(1, (0, 0, 0))) == rgb_mix     # child at 8,0,0
    assert albedo_of(t.sample_root_mip(OOB_OCTANT, (0, 0, 0))) == rg_mix
    assert albedo_of(t.sample_root_mip(OOB_OCTANT, (2, 0, 0))) == rgb_mix


# tests.rs:365-468
def test_mixed_mip_lvl2_where_dim_is_4(Tree):
    t = Tree(16, 4)
    t.set_auto_simplify(False)
    t.switch_albedo_mip_maps(True).set_method_at(1, BOX).set_method_at(2, BOX)
    _dim4_scene(t)
    _dim4_asserts(t)


# tests.rs:471-567
def test_mixed_mip_regeneration_lvl2_where_dim_is_4(Tree):
    t = Tree(16, 4)
    t.set_auto_simplify(False)
    _dim4_scene(t)
    t.switch_albedo_mip_maps(True).set_method_at(1, BOX).set_method_at(2, BOX).recalculate_mips()
    _dim4_asserts(t)


# mipmap.rs:591-604, :610-672, :718-721
def test_strategy_defaults_and_setters(Tree):
    t = Tree(16, 4)
    assert not t.mip_enabled()
    assert t.get_method_at(1)[0] == POINT
    for lvl in (2, 3, 4, 7):
        assert t.get_method_at(lvl)[0] == BOX
    assert t.get_new_color_similarity_at(1) == 0.0
    assert t.get_new_color_similarity_at(2) == f32(0.1)
    assert t.get_new_color_similarity_at(3) == f32(0.05)
    assert t.get_new_color_similarity_at(4) == f32(0.02)
    t.set_method_at(3, POSTERIZE, 1.5)   # clamped to [0, 1] (mipmap.rs:662-670)
    assert t.get_method_at(3) == (POSTERIZE, 1.0)
    t.set_method_at(2, POSTERIZE_BD, 0.25)
    assert t.get_method_at(2) == (POSTERIZE_BD, 0.25)
    t.set_color_similarity_thr_at(5, 2.0)
    assert t.get_new_color_similarity_at(5) == 1.0
    t.switch_albedo_mip_maps(True)
    assert t.mip_enabled()
    t.mip_reset()  # back to MIPMapStrategy::default(): disabled
    assert not t.mip_enabled()
    assert t.get_method_at(3)[0] == BOX
    assert t.get_new_color_similarity_at(5) == 0.0


def test_uniform_leaf_has_no_mip_and_point_filter_reuses_colours(Tree):
    # mipmap.rs:331-337: a UniformLeaf's content is its own MIP
    t = Tree(8, 2)
    t.switch_albedo_mip_maps(True)
    t.insert_at_lod((0, 0, 0), 4, RED)
    assert t.sample_root_mip(0, (0, 0, 0)) == (O.EMPTY,)
    # level 1 default = PointFilter: the most frequent colour of the 2x2x2 cell, no new palette entry
    t2 = Tree(8, 2)
    t2.set_auto_simplify(False)
    t2.switch_albedo_mip_maps(True)
    for p, c in [((0, 0, 0), RED), ((0, 0, 1), GREEN), ((0, 1, 0), GREEN), ((1, 1, 1), GREEN), ((1, 0, 0), BLUE)]:
        t2.insert(p, c)
    assert albedo_of(t2.sample_root_mip(0, (0, 0, 0))) == (0, 255, 0, 255)


def _random_ops(rng, size, n):
    ops = []
    for _ in range(n):
        k = rng.integers(0, 10)
        pos = tuple(int(v) for v in rng.integers(0, size, 3))
        col = int(rng.choice([RED, GREEN, BLUE, 0x808080FF, 0x10F0A0FF, 0xFFFFFFFF, 0x7F7F0080]))
        if k < 6:
            ops.append(("insert", pos, col))
        elif k < 8:
            ops.append(("insert_at_lod", pos, int(rng.choice([2, 4, 8])), col))
        elif k < 9:
            ops.append(("clear", pos))
        else:
            ops.append(("clear_at_lod", pos, int(rng.choice([2, 4]))))
    return ops


def _apply(t, ops):
    for op in ops:
        if op[0] == "insert":
            t.insert(op[1], op[2])
        elif op[0] == "insert_at_lod":
            t.insert_at_lod(op[1], op[2], op[3])
        elif op[0] == "clear":
            t.clear(op[1])
        else:
            t.clear_at_lod(op[1], op[2])


@pytest.mark.parametrize("size,dim", [(8, 1), (16, 2), (32, 4), (64, 8)])
@pytest.mark.parametrize("methods", [
    {},                                             # defaults: PointFilter on 1, BoxFilter above, thresholds on 2..4
    {1: (BOX, 0.0), 2: (POINT, 0.0)},
    {1: (POINT_BD, 0.0), 2: (POINT_BD, 0.0), 3: (POINT_BD, 0.0)},
    {1: (POSTERIZE, 0.2), 2: (POSTERIZE, 0.1), 3: (POSTERIZE_BD, 0.3)},
])
def test_product_mips_equal_oracle_mips_under_random_edits(size, dim, methods):
    """Two independent implementations (oracle: owned Vec bricks; product: pooled bricks) must agree on every MIP
    brick, on the palette the MIPs extended, and on the tree itself, while edits update the MIPs incrementally
    (insert.rs:371, clear.rs:335) and after a full recalculate_mips()."""
    rng = np.random.default_rng(size * 131 + dim + 7 * len(methods))
    for simplify in (True, False):
        a, b = OracleOctree(size, dim), ProductOctree(size, dim)
        for t in (a, b):
            t.set_auto_simplify(simplify)
            t.switch_albedo_mip_maps(True)
            for lvl, (m, thr) in methods.items():
                t.set_method_at(lvl, m, thr)
        for round_ in range(3):
            ops = _random_ops(rng, size, 150)
            _apply(a, ops)
            _apply(b, ops)
            assert a.structure_hash() == b.structure_hash()
            assert a.mip_hash() == b.mip_hash(), (size, dim, methods, simplify, round_)
        a.recalculate_mips()
        b.recalculate_mips()
        assert a.structure_hash() == b.structure_hash()
        assert a.mip_hash() == b.mip_hash()


def test_enabling_mips_later_equals_recalculation():
    rng = np.random.default_rng(5)
    ops = _random_ops(rng, 32, 300)
    for cls in (OracleOctree, ProductOctree):
        t = cls(32, 4)
        _apply(t, ops)
        t.switch_albedo_mip_maps(True)   # recalculates (mipmap.rs:866-871)
        h = t.mip_hash()
        t.recalculate_mips()
        # BoxFilter may have added palette colours the first time; the bricks themselves must not move
        assert t.mip_hash() == h


def test_nodes_smaller_than_a_brick_keep_their_mips_untouched():
    """insert_at_lod with a size <= brick_dim splits leaves into nodes smaller than one brick (insert.rs:137-147). The
    reference's update_mip then computes `position % (size / dim)` = `% 0` and panics; oracle and product skip the MIP
    update of such a node instead (and must agree with each other everywhere else)."""
    a, b = OracleOctree(16, 4), ProductOctree(16, 4)
    for t in (a, b):
        t.switch_albedo_mip_maps(True)
        t.insert((9, 9, 13), RED)
        t.insert_at_lod((8, 8, 12), 2, GREEN)      # aligned and lexicographically <= the child's corner
        t.insert_at_lod((4, 6, 2), 2, BLUE)
        t.insert((8, 9, 13), 0x123456FF)
        t.insert((5, 6, 2), 0x654321FF)
        t.clear((8, 8, 12))
    assert a.structure_hash() == b.structure_hash()
    assert a.mip_hash() == b.mip_hash()
    a.recalculate_mips()
    b.recalculate_mips()
    assert a.mip_hash() == b.mip_hash()
