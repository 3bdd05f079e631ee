"""Black-box construction tests of the reference (src/octree/update/tests.rs, src/octree/mod.rs:173-205),
restated against the CPU oracle AND the product's host octree. Paths relative to /root/reference/."""
import itertools

import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OracleOctree, entry_key as K
from product_adapter import ProductOctree

RED, GREEN, BLUE = 0xFF0000FF, 0x00FF00FF, 0x0000FFFF
OFFS = [(0, 0, 0), (1, 0, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (1, 1, 0), (0, 1, 1), (1, 1, 1)]  # lut.rs:156-197


def make_tree(cls, size, dim, simplify=True):
    t = cls(size, dim)
    if not simplify:
        t.set_auto_simplify(False)
    return t


@pytest.fixture(params=["oracle", "product"])
def Tree(request):
    """Every construction test runs against the CPU oracle and against the product's host octree (C ABI)."""
    return OracleOctree if request.param == "oracle" else ProductOctree


# src/octree/mod.rs:173-187 (validation order)
def test_new_validation(Tree):
    def status(size, dim):
        try:
            Tree(size, dim)
            return O.OK
        except ValueError as e:
            return e.args[0]

    assert status(64, 8) == O.OK
    assert status(2, 1) == O.OK
    assert status(1024, 32) == O.OK
    assert status(0, 8) == O.E_INVALID_BRICK_DIMENSION
    assert status(64, 3) == O.E_INVALID_BRICK_DIMENSION
    assert status(4, 8) == O.E_INVALID_SIZE
    assert status(24, 8) == O.E_INVALID_SIZE
    assert status(8, 8) == O.E_INVALID_STRUCTURE
    for k in range(0, 12):  # every power-of-two brick dimension the examples could use
        assert status(2 << k, 1 << k) == O.OK


# update/tests.rs:9-17
def test_simplest_insert_and_get(Tree):
    t = make_tree(Tree, 2, 1, False)
    assert t.insert((0, 0, 0), 0xFF000001) == O.OK
    assert t.get((0, 0, 0)) == K(0xFF000001)


# update/tests.rs:20-46
def test_simple_insert_and_get(Tree):
    t = make_tree(Tree, 2, 1, False)
    t.insert((1, 0, 0), RED)
    t.insert((0, 1, 0), GREEN)
    t.insert((0, 0, 1), BLUE)
    assert t.get((1, 0, 0)) == K(RED) and t.get((0, 1, 0)) == K(GREEN) and t.get((0, 0, 1)) == K(BLUE)
    assert t.get((1, 1, 1)) == K()
    t.insert((1, 0, 0), GREEN)
    assert t.get((1, 0, 0)) == K(GREEN) and t.get((0, 1, 0)) == K(GREEN) and t.get((0, 0, 1)) == K(BLUE)
    assert t.get((1, 1, 1)) == K()


# update/tests.rs:49-54, :1676-1687
def test_insert_empty(Tree):
    t = make_tree(Tree, 2, 1, False)
    assert t.insert((0, 0, 0)) == O.OK
    assert t.get((0, 0, 0)) == K()
    t = make_tree(Tree, 4, 1)
    assert t.insert((3, 0, 0), (0, 0, 0, 0)) == O.OK
    assert t.get((3, 0, 0)) == K()


def test_insert_out_of_bounds(Tree):
    t = make_tree(Tree, 4, 1)
    assert t.insert((4, 0, 0), RED) == O.E_INVALID_POSITION


# update/tests.rs:57-110
def test_complex_insert_and_get(Tree):
    t = make_tree(Tree, 2, 1, False)
    t.insert((1, 0, 0), RED, 3)
    t.insert((0, 1, 0), GREEN, 1)
    t.insert((0, 0, 1), None, 2)
    assert t.get((1, 0, 0)) == K(RED, 3)
    assert t.get((0, 1, 0)) == K(GREEN, 1)
    assert t.get((0, 0, 1)) == K(None, 2)
    assert t.get((1, 1, 1)) == K()
    t.insert((1, 0, 0), None, 3)
    assert t.get((1, 0, 0)) == K(None, 3)
    assert t.get((0, 1, 0)) == K(GREEN, 1)
    assert t.get((0, 0, 1)) == K(None, 2)
    assert t.get((1, 1, 1)) == K()


# update/tests.rs:113-140
def test_simple_insert_and_get_where_dim_is_2(Tree):
    t = make_tree(Tree, 4, 2, False)
    t.insert((1, 0, 0), RED); t.insert((0, 1, 0), GREEN); t.insert((0, 0, 1), BLUE)
    assert t.get((1, 0, 0)) == K(RED) and t.get((0, 1, 0)) == K(GREEN) and t.get((0, 0, 1)) == K(BLUE)
    t.insert((3, 0, 0), RED); t.insert((0, 3, 0), GREEN); t.insert((0, 0, 3), BLUE)
    assert t.get((3, 0, 0)) == K(RED) and t.get((0, 3, 0)) == K(GREEN) and t.get((0, 0, 3)) == K(BLUE)
    t.insert((1, 0, 0), GREEN)
    assert t.get((1, 0, 0)) == K(GREEN) and t.get((0, 1, 0)) == K(GREEN) and t.get((0, 0, 1)) == K(BLUE)


def count_hits(t, n, want):
    hits = 0
    for p in itertools.product(range(n), repeat=3):
        g = t.get(p)
        if g != K():
            assert g == K(want), p
            hits += 1
    return hits


# update/tests.rs:143-191 and :194-242
@pytest.mark.parametrize("dim", [1, 2])
def test_insert_at_lod(Tree, dim):
    t = make_tree(Tree, 8, dim, False)
    t.insert_at_lod((0, 0, 0), 2, RED)
    for p in itertools.product(range(2), repeat=3):
        assert t.get(p) == K(RED)
    t.insert_at_lod((0, 0, 0), 4, GREEN)
    assert count_hits(t, 4, GREEN) == 64


# update/tests.rs:361-401
def test_update(Tree):
    t = make_tree(Tree, 2, 1, False)
    t.insert((0, 0, 0), RED, 3)
    assert t.get((0, 0, 0)) == K(RED, 3)
    t.update((0, 0, 0), GREEN)
    assert t.get((0, 0, 0)) == K(GREEN, 3)
    t = make_tree(Tree, 2, 1, False)
    t.insert((0, 0, 0), RED, 3)
    t.update((0, 0, 0), None, 4)
    assert t.get((0, 0, 0)) == K(RED, 4)
    t.update((0, 0, 0))
    assert t.get((0, 0, 0)) == K(RED, 4)


# update/tests.rs:432-460
def test_uniform_solid_leaf_separated_by_insert_where_dim_is_1(Tree):
    t = make_tree(Tree, 2, 1)
    for o in OFFS:
        t.insert(o, 0xFFFF00FF)
    assert t.get((0, 0, 0)) == K(0xFFFF00FF)
    t.insert((0, 0, 0), 0xFFFF00FF)
    for o in OFFS:
        assert t.get(o) == K(0xFFFF00FF)


# update/tests.rs:515-567
def test_uniform_solid_leaf_separated_by_insert_where_dim_is_4(Tree):
    D = 4
    t = make_tree(Tree, 8, D)
    base = 0xFFFF00AA
    for octant, o in enumerate(OFFS):
        start = [c * (D // 2) for c in o]
        for d in itertools.product(range(D // 2), repeat=3):
            t.insert(tuple(s + e for s, e in zip(start, d)), base + octant)
    assert t.get((0, 0, 0)) == K(base)
    t.insert((0, 0, 0), 0x000000FF)
    assert t.get((0, 0, 0)) == K(0x000000FF)
    for octant, o in enumerate(OFFS):
        start = [c * (D // 2) for c in o]
        for d in itertools.product(range(D // 2), repeat=3):
            if d == (0, 0, 0) and octant == 0:
                continue
            assert t.get(tuple(s + e for s, e in zip(start, d))) == K(base + octant)


# update/tests.rs:643-692
def test_simple_uniform_parted_brick_leaf_separated_by_insert(Tree):
    D = 2
    t = make_tree(Tree, 4, D)
    base0 = 0xF00000FF
    for octant, o in enumerate(OFFS):
        t.insert_at_lod(tuple(c * D for c in o), D, base0 + 2 * octant)
    assert t.get((0, 0, 0)) == K(base0)
    t.insert((0, 0, 0), 0x000000FF)
    assert t.get((0, 0, 0)) == K(0x000000FF)
    for d in itertools.product(range(D), repeat=3):
        for octant, o in enumerate(OFFS):
            p = tuple(c * D + e for c, e in zip(o, d))
            if d == (0, 0, 0) and octant == 0:
                assert t.get(p) == K(0x000000FF)
            else:
                assert t.get(p) == K(base0 + 2 * octant)


# update/tests.rs:695-771
def test_uniform_parted_brick_leaf_separated_by_insert_where_dim_is_4(Tree):
    D = 4
    t = make_tree(Tree, 8, D)

    def color(x, y, z):
        return (x - x % (D // 2), y - y % (D // 2), z - z % (D // 2), 255)

    for o in OFFS:
        for x, y, z in itertools.product(range(D), repeat=3):
            p = (o[0] * D + x, o[1] * D + y, o[2] * D + z)
            t.insert(p, color(x, y, z))
            assert t.get(p) == K(color(x, y, z)), p
    assert t.get((0, 0, 0)) == K(0x000000FF)
    t.insert((1, 1, 1), 0xFF0000FF)
    assert t.get((1, 1, 1)) == K(0xFF0000FF)
    assert t.get((0, 0, 0)) == K(0x000000FF)
    for octant, o in enumerate(OFFS):
        for x, y, z in itertools.product(range(D), repeat=3):
            p = (o[0] * D + x, o[1] * D + y, o[2] * D + z)
            if (x, y, z) == (1, 1, 1) and octant == 0:
                assert t.get(p) == K(0xFF0000FF)
            else:
                assert t.get(p) == K(color(x, y, z)), p


# update/tests.rs:774-909
def test_insert_at_lod_unaligned(Tree):
    t = make_tree(Tree, 8, 4, False)
    t.insert_at_lod((1, 1, 1), 4, RED)
    assert count_hits(t, 4, RED) == 27
    t = make_tree(Tree, 8, 1, False)
    t.insert_at_lod((2, 2, 2), 3, RED)
    assert count_hits(t, 8, RED) == 8
    t = make_tree(Tree, 8, 1, False)
    t.insert_at_lod((3, 3, 3), 3, RED)
    assert count_hits(t, 8, RED) == 1
    t = make_tree(Tree, 8, 4, False)
    t.insert_at_lod((1, 1, 1), 3, RED)
    assert t.get((1, 1, 1)) == K(RED)
    assert count_hits(t, 8, RED) == 27


# update/tests.rs:912-965
def test_insert_at_lod_with_simplify(Tree):
    t = make_tree(Tree, 8, 1)
    t.insert_at_lod((4, 0, 0), 2, RED)
    for x, y, z in itertools.product((4, 5), (0, 1), (0, 1)):
        assert t.get((x, y, z)) == K(RED)
    t.insert_at_lod((0, 0, 0), 4, GREEN)
    hits = 0
    for p in itertools.product(range(4), repeat=3):
        g = t.get(p)
        if g != K():
            assert g == K(GREEN)
            hits += 1
    for p in itertools.product((4, 5), (0, 1), (0, 1)):
        g = t.get(p)
        if g != K():
            assert g == K(RED)
            hits += 1
    assert hits == 64 + 8


# update/tests.rs:968-1027
@pytest.mark.parametrize("size,dim", [(2, 1), (4, 2)])
def test_simplifyable_insert_and_get(Tree, size, dim):
    t = make_tree(Tree, size, dim)
    for p in itertools.product(range(size), repeat=3):
        t.insert(p, RED)
    t.insert((0, 0, 0), GREEN)
    assert t.get((0, 0, 0)) == K(GREEN)
    for p in itertools.product(range(1, size), repeat=3):
        assert t.get(p) == K(RED)


# update/tests.rs:1088-1118 (insert halves; clear() is out of scope)
def test_set_small_part_of_large_node(Tree):
    t = make_tree(Tree, 64, 8)
    t.insert_at_lod((33, 33, 33), 2, RED)
    assert t.get((33, 33, 33)) == K(RED)
    t = make_tree(Tree, 64, 8)
    t.insert((31, 31, 31), RED)
    assert t.get((31, 31, 31)) == K(RED)


# update/tests.rs:1601-1655
def test_overwrite_whole_nodes_where_dim_is_4(Tree):
    t = make_tree(Tree, 16, 4)
    t.insert_at_lod((0, 0, 0), 8, RED)
    assert count_hits(t, 8, RED) == 512
    t.insert_at_lod((0, 0, 0), 5, BLUE)
    reds = blues = 0
    for p in itertools.product(range(8), repeat=3):
        g = t.get(p)
        assert g != K()
        reds += g == K(RED)
        blues += g == K(BLUE)
    assert reds == 512 - 64 and blues == 64


# update/tests.rs:1658-1673
def test_edge_case_octree_set(Tree):
    t = make_tree(Tree, 8, 1)
    for p in itertools.product(range(6, 8), repeat=3):
        t.insert(p, sum(p))
        assert t.get(p) == K(sum(p))


# a randomised model check: every get() equals a dict model after arbitrary inserts (seeded)
@pytest.mark.parametrize("size,dim", [(8, 1), (8, 2), (16, 4), (32, 8)])
def test_random_inserts_match_model(Tree, size, dim):
    rng = np.random.default_rng(size * 100 + dim)
    t = make_tree(Tree, size, dim)
    model = {}
    colors = [0xFF0000FF, 0x00FF00FF, 0x0000FFFF, 0x112233FF]
    for _ in range(600):
        p = tuple(int(v) for v in rng.integers(0, size, 3))
        c = colors[int(rng.integers(0, len(colors)))]
        assert t.insert(p, c) == O.OK
        model[p] = c
    for p in itertools.product(range(size), repeat=3):
        assert t.get(p) == (K(model[p]) if p in model else K()), p


# ---- clear / clear_at_lod (src/octree/update/clear.rs), tests of src/octree/update/tests.rs -------------------------
def fill(t, size, color):
    for p in itertools.product(range(size), repeat=3):
        t.insert(p, color)


# update/tests.rs:245-358
@pytest.mark.parametrize("dim", [1, 2, 4])
def test_case_simplified_insert_separated_by_clear(Tree, dim):
    t = make_tree(Tree, 8, dim)
    fill(t, 8, RED)
    assert t.get((3, 3, 3)) == K(RED)
    assert t.clear((3, 3, 3)) == O.OK
    assert t.get((3, 3, 3)) == K()
    assert count_hits(t, 8, RED) == 511


# update/tests.rs:404-429
def test_uniform_solid_leaf_separated_by_clear_where_dim_is_1(Tree):
    t = make_tree(Tree, 2, 1)
    for o in OFFS:
        t.insert(o, 0xFFFF00FF)
    t.clear((0, 0, 0))
    assert t.get((0, 0, 0)) == K()
    for o in OFFS[1:]:
        assert t.get(o) == K(0xFFFF00FF)


# update/tests.rs:463-512
def test_uniform_solid_leaf_separated_by_clear_where_dim_is_4(Tree):
    D = 4
    t = make_tree(Tree, 8, D)
    base = 0xFFFF00AA
    for octant, o in enumerate(OFFS):
        start = [c * (D // 2) for c in o]
        for d in itertools.product(range(D // 2), repeat=3):
            t.insert(tuple(s + e for s, e in zip(start, d)), base + octant)
    assert t.get((0, 0, 0)) == K(base)
    t.clear((0, 0, 0))
    assert t.get((0, 0, 0)) == K()
    for octant, o in enumerate(OFFS):
        start = [c * (D // 2) for c in o]
        for d in itertools.product(range(D // 2), repeat=3):
            if d == (0, 0, 0) and octant == 0:
                continue
            assert t.get(tuple(s + e for s, e in zip(start, d))) == K(base + octant)


# update/tests.rs:570-640
def test_uniform_parted_brick_leaf_separated_by_clear_where_dim_is_4(Tree):
    D = 4
    t = make_tree(Tree, 8, D)

    def color(x, y, z):
        return (x - x % (D // 2), y - y % (D // 2), z - z % (D // 2), 255)

    for o in OFFS:
        for x, y, z in itertools.product(range(D), repeat=3):
            t.insert((o[0] * D + x, o[1] * D + y, o[2] * D + z), color(x, y, z))
    assert t.get((0, 0, 0)) == K(0x000000FF)
    t.clear((1, 1, 1))
    assert t.get((1, 1, 1)) == K()
    for octant, o in enumerate(OFFS):
        for x, y, z in itertools.product(range(D), repeat=3):
            p = (o[0] * D + x, o[1] * D + y, o[2] * D + z)
            if (x, y, z) == (1, 1, 1) and octant == 0:
                assert t.get(p) == K()
            else:
                assert t.get(p) == K(color(x, y, z)), p


# update/tests.rs:1030-1069
@pytest.mark.parametrize("size,dim", [(2, 1), (4, 2)])
def test_simple_clear(Tree, size, dim):
    t = make_tree(Tree, size, dim, False)
    t.insert((1, 0, 0), RED); t.insert((0, 1, 0), GREEN); t.insert((0, 0, 1), BLUE)
    t.clear((0, 0, 1))
    assert t.get((1, 0, 0)) == K(RED) and t.get((0, 1, 0)) == K(GREEN)
    assert t.get((0, 0, 1)) == K() and t.get((1, 1, 1)) == K()


# update/tests.rs:1072-1118
def test_clear_small_parts_of_large_nodes(Tree):
    t = make_tree(Tree, 64, 8)
    t.insert((0, 1, 1), RED); t.insert((1, 0, 0), RED)
    t.clear((1, 0, 0))
    assert t.get((1, 0, 0)) == K() and t.get((0, 1, 1)) == K(RED)
    t = make_tree(Tree, 64, 8)
    t.insert_at_lod((33, 33, 33), 2, RED)
    assert t.get((33, 33, 33)) == K(RED)
    t.clear((33, 33, 33))
    assert t.get((33, 33, 33)) == K()
    t = make_tree(Tree, 64, 8)
    t.insert((31, 31, 31), RED)
    t.clear((31, 31, 31))
    assert t.get((31, 31, 31)) == K()


# update/tests.rs:1121-1136
def test_double_clear(Tree):
    t = make_tree(Tree, 2, 1, False)
    t.insert((1, 0, 0), 0x000000FF); t.insert((0, 1, 0), 0xFFFFFFFF); t.insert((0, 0, 1), 0xFFFFFFFF)
    t.clear((0, 0, 1)); t.clear((0, 0, 1))
    assert t.get((1, 0, 0)) == K(0x000000FF) and t.get((0, 1, 0)) == K(0xFFFFFFFF) and t.get((0, 0, 1)) == K()


# update/tests.rs:1139-1196
@pytest.mark.parametrize("size,dim", [(2, 1), (4, 2)])
def test_simplifyable_clear(Tree, size, dim):
    t = make_tree(Tree, size, dim)
    fill(t, size, 0xFFAAEEFF)
    t.clear((0, 0, 0))
    assert t.get((0, 0, 0)) == K()
    for p in itertools.product(range(1, size), repeat=3):
        assert t.get(p) == K(0xFFAAEEFF)


# update/tests.rs:1199-1226
def test_clear_to_nothing(Tree):
    t = make_tree(Tree, 4, 1)
    for p in itertools.product(range(2), repeat=3):
        t.insert(p, 0xFFAAEEFF)
    t.clear_at_lod((0, 0, 0), 2)
    for p in itertools.product(range(2), repeat=3):
        assert t.get(p) == K()


# update/tests.rs:1229-1278
def test_clear_edge_case(Tree):
    t = make_tree(Tree, 64, 16)
    t.update((1, 0, 0), None, 0xFACEFEED)
    t.insert_at_lod((0, 0, 0), 32, RED)
    t.clear_at_lod((5, 5, 5), 8)
    for p in itertools.product(range(5, 8), repeat=3):
        assert t.get(p) == K()
    for p in itertools.product(range(5), repeat=3):
        assert t.get(p) == K(RED), p
    t.clear_at_lod((0, 0, 0), 32)
    for p in itertools.product(range(0, 32, 3), repeat=3):
        assert t.get(p) == K(), p


# update/tests.rs:1281-1350
@pytest.mark.parametrize("dim", [1, 2])
def test_clear_at_lod(Tree, dim):
    t = make_tree(Tree, 8, dim)
    t.insert_at_lod((0, 0, 0), 4, 0xFFAAEEFF)
    t.clear_at_lod((0, 0, 0), 2)
    assert count_hits(t, 4, 0xFFAAEEFF) == 64 - 8


# update/tests.rs:1353-1405
def test_clear_at_lod_with_unaligned_position(Tree):
    t = make_tree(Tree, 8, 1)
    t.insert_at_lod((0, 0, 0), 4, 0xFFAAEEFF)
    t.clear_at_lod((1, 1, 1), 2)
    for p in itertools.product(range(2), repeat=3):
        assert t.get(p) == K()
    for p in [(0, 0, 2), (0, 2, 0), (0, 2, 2), (2, 0, 0), (2, 0, 2), (2, 2, 0), (2, 2, 2)]:
        assert t.get(p) != K()
    assert count_hits(t, 4, 0xFFAAEEFF) == 64 - 8


# update/tests.rs:1408-1463
def test_clear_at_lod_with_unaligned_position_where_dim_is_4(Tree):
    t = make_tree(Tree, 16, 4)
    t.insert_at_lod((0, 0, 0), 8, 0xFFAAEEFF)
    assert count_hits(t, 8, 0xFFAAEEFF) == 512
    t.clear_at_lod((1, 1, 1), 4)
    assert count_hits(t, 8, 0xFFAAEEFF) == 512 - 27


# update/tests.rs:1466-1547
def test_clear_at_lod_with_unaligned_size(Tree):
    t = make_tree(Tree, 8, 1)
    t.insert_at_lod((0, 0, 0), 4, 0xFFAAEEFF)
    t.clear_at_lod((0, 0, 0), 3)
    assert count_hits(t, 4, 0xFFAAEEFF) == 64 - 8
    t = make_tree(Tree, 8, 4)
    t.insert_at_lod((0, 0, 0), 4, 0xFFAAEEFF)
    assert count_hits(t, 8, 0xFFAAEEFF) == 64
    t.clear_at_lod((0, 0, 0), 3)
    assert count_hits(t, 8, 0xFFAAEEFF) == 64 - 27


# update/tests.rs:1550-1598
def test_clear_whole_nodes_where_dim_is_4(Tree):
    t = make_tree(Tree, 16, 4)
    t.insert_at_lod((0, 0, 0), 8, 0xFFAAEEFF)
    assert count_hits(t, 8, 0xFFAAEEFF) == 512
    t.clear_at_lod((0, 0, 0), 5)
    assert count_hits(t, 8, 0xFFAAEEFF) == 512 - 64


def test_clear_out_of_bounds(Tree):
    t = make_tree(Tree, 4, 1)
    assert t.clear((0, 4, 0)) == O.E_INVALID_POSITION


@pytest.mark.parametrize("size,dim", [(8, 2), (16, 4), (32, 8)])
def test_random_inserts_and_clears_match_model(Tree, size, dim):
    """Random inserts and clears against a dict model. brick_dim 1 is left out on purpose: there the reference turns
    Internal nodes of size 2 back into leaves and drops their children (update/mod.rs:497-521, "might induce data loss -
    see #69"), which both implementations reproduce (tests/test_host_octree_shape.py checks they agree)."""
    rng = np.random.default_rng(size * 1000 + dim)
    t = make_tree(Tree, size, dim)
    model = {}
    colors = [0xFF0000FF, 0x00FF00FF, 0x0000FFFF]
    for step in range(900):
        p = tuple(int(v) for v in rng.integers(0, size, 3))
        if rng.integers(0, 3) == 0 and model:
            if rng.integers(0, 2):
                p = list(model)[int(rng.integers(0, len(model)))]
            assert t.clear(p) == O.OK
            model.pop(p, None)
        else:
            c = colors[int(rng.integers(0, len(colors)))]
            assert t.insert(p, c) == O.OK
            model[p] = c
    for p in itertools.product(range(size), repeat=3):
        assert t.get(p) == (K(model[p]) if p in model else K()), p


# voxel_color_palette / voxel_data_palette (src/octree/types.rs:191-192, add_to_palette in detail.rs): first-seen order,
# no duplicates - what svx_octree_color_palette / svx_octree_data_palette hand out
def test_palettes_are_readable_in_first_seen_order():
    p, o = ProductOctree(16, 4), OracleOctree(16, 4)
    for t in (p, o):
        t.insert((0, 0, 0), RED)
        t.insert((1, 0, 0), GREEN, 7)
        t.insert((2, 0, 0), RED, 9)
        t.insert((3, 0, 0), None, 7)
        t.insert((4, 4, 4), BLUE)
        t.insert((5, 5, 5), GREEN)
    colors, data = p.tree.color_palette(), p.tree.data_palette()
    assert colors.tolist() == [[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255]]
    assert data.tolist() == [7, 9]
    n_data = O.C.c_uint64(0)
    assert O.lib().svxo_octree_palette_sizes(o._h, O.C.byref(n_data)) == len(colors) and n_data.value == len(data)
    empty = ProductOctree(8, 2)
    assert empty.tree.color_palette().shape == (0, 4) and empty.tree.data_palette().shape == (0,)
