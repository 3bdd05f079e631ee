"""The reference's remaining unit-level known-answer tests, run against the oracle's restatement of the same
internals: ObjectPool (src/object_pool.rs:243-285: the allocation order decides node keys, and with them the child
links every ray follows), BrickData::is_empty_throughout / is_part_empty_throughout (src/octree/tests.rs:8-151:
`clear` decides with them which children to drop) and the per-type bencode round trips
(src/convert/bytecode_tests.rs:7-180) with the reference's own values. Paths relative to /root/reference/.
"""
import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OracleOctree
from test_bytecode import bdec, benc, tree_doc

L = O.lib()
NIL = 0xFFFFFFFF
NONE = np.iinfo(np.int64).min
PUSH, POP, FREE, GET, SET, FIRST_AVAILABLE, LEN, VALID = range(8)


def _pool(ops):
    a = np.array(ops, dtype=np.int64).reshape(-1, 2)
    out = np.zeros(len(a), dtype=np.int64)
    L.svxo_node_pool_script(a.ctypes.data, len(a), out.ctypes.data)
    return [None if v == NONE else int(v) for v in out]


# src/object_pool.rs:247-259
def test_push_pop_modify():
    r = _pool([(PUSH, 5), (GET, 0), (SET, 0 | (10 << 32)), (GET, 0), (POP, 0), (POP, 0)])
    assert r == [0, 5, 0, 10, 10, None]


# src/object_pool.rs:261-270
def test_push_deallocate():
    r = _pool([(PUSH, 5), (GET, 0), (FREE, 0), (POP, 0), (FREE, 0)])
    assert r == [0, 5, 1, None, 0]


# src/object_pool.rs:272-284
def test_edge_case_reused_item():
    r = _pool([(PUSH, 5), (PUSH, 10), (POP, 0), (FIRST_AVAILABLE, 0), (PUSH, 15), (GET, 0), (LEN, 0)])
    assert r == [0, 1, 5, 0, 0, 15, 2]  # the original key is reused to hold the latest value


def test_pool_hands_out_the_lowest_run_of_free_keys_first():
    """allocate() (object_pool.rs:178-202): `first_available` only ever looks one slot ahead, so after freeing keys
    1, 2 and 4 of six the next pushes get 1, 2 - and then a NEW key 6, not 4: the cursor jumped to the end of the
    buffer when slot 3 turned out reserved (check_first_available, :156-166). Key 4 is handed out only after
    another free moves the cursor back. Node keys - and so the saved file - depend on exactly this."""
    r = _pool([(PUSH, i) for i in range(6)] + [(FREE, 1), (FREE, 2), (FREE, 4), (FIRST_AVAILABLE, 0),
              (PUSH, 100), (FIRST_AVAILABLE, 0), (PUSH, 101), (FIRST_AVAILABLE, 0), (PUSH, 102), (LEN, 0),
              (VALID, 4), (FREE, 0), (PUSH, 103), (PUSH, 104), (LEN, 0)])
    assert r[:6] == list(range(6))
    assert r[6:] == [1, 1, 1, 1, 1, 2, 2, 2, 6, 7, 0, 1, 0, 7, 8]


def _empty(dim, kind, voxels, target, part=-1, colors=(0x00000064,), datas=(0,)):
    v = np.array(voxels if len(voxels) else [0], dtype=np.uint32)
    c = np.array(colors, dtype=np.uint32)
    d = np.array(datas, dtype=np.uint32)
    r = L.svxo_brick_is_empty_throughout(dim, kind, v.ctypes.data, len(voxels), part, target, c.ctypes.data, len(c), d.ctypes.data, len(d))
    assert r in (0, 1)
    return bool(r)


EMPTY_BRICK, PARTED, SOLID = 0, 1, 2
VISUAL_0 = 0xFFFF0000  # NodeContent::pix_visual(0): colour index 0, no data index (node.rs:354-403)
OFFSET = [(0, 0, 0), (1, 0, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (1, 1, 0), (0, 1, 1), (1, 1, 1)]  # OCTANT_OFFSET_REGION_LUT


def flat(x, y, z, dim):
    return x + y * dim + z * dim * dim


# src/octree/tests.rs:8-31 ; the palette colour is Albedo::default().with_alpha(100)
def test_octant_empty():
    for o in range(8):
        assert _empty(1, EMPTY_BRICK, [], o)
        assert not _empty(1, PARTED, [VISUAL_0], o)


# src/octree/tests.rs:33-62
def test_octant_empty_where_dim_is_2():
    brick = [VISUAL_0] * 8
    for o in range(8):
        assert not _empty(2, PARTED, brick, o)
    brick[flat(*OFFSET[5], 2)] = NIL
    assert _empty(2, PARTED, brick, 5), "Data cleared under octant should be empty"
    assert [o for o in range(8) if _empty(2, PARTED, brick, o)] == [5]


# src/octree/tests.rs:64-117
def test_octant_empty_where_dim_is_4():
    brick = [VISUAL_0] * 64
    for o in range(8):
        assert not _empty(4, PARTED, brick, o)
    ox, oy, oz = (2 * v for v in OFFSET[5])
    brick[flat(ox, oy, oz, 4)] = NIL
    assert not _empty(4, PARTED, brick, 5), "Data cleared under octant should not be empty"
    for x in range(2):
        for y in range(2):
            for z in range(2):
                brick[flat(ox + x, oy + y, oz + z, 4)] = NIL
    assert _empty(4, PARTED, brick, 5), "Data cleared under octant should be empty"
    assert [o for o in range(8) if _empty(4, PARTED, brick, o)] == [5]


# src/octree/tests.rs:119-151
def test_part_of_octant_empty():
    brick = [VISUAL_0] * 64
    for i in range(8):
        for j in range(8):
            assert not _empty(4, PARTED, brick, j, part=i)
    ox, oy, oz = (2 * v for v in OFFSET[5])
    brick[flat(ox, oy, oz, 4)] = NIL
    assert _empty(4, PARTED, brick, 0, part=5), "Data cleared under part of octant should be empty"
    assert not _empty(4, PARTED, brick, 1, part=5), "Data not cleared should not be empty"


def test_solid_and_transparent_bricks_are_empty_by_their_palette_entry():
    """node.rs:113-118, :190-195: a Solid brick is empty iff its value points to nothing - an absent index or a colour
    with alpha 0 (pix_points_to_empty, :405-427), a data entry that is zero."""
    assert not _empty(2, SOLID, [VISUAL_0], 3)
    assert _empty(2, SOLID, [NIL], 3)
    assert _empty(2, SOLID, [VISUAL_0], 3, colors=(0x11223300,))          # alpha 0
    assert _empty(2, SOLID, [0x0000FFFF], 3, datas=(0,))                  # data 0 = u32::zero() is empty
    assert not _empty(2, SOLID, [0x0000FFFF], 3, datas=(7,))
    assert not _empty(2, SOLID, [0x00000000], 3, colors=(0x11223300,), datas=(7,), part=2)


# ---- src/convert/bytecode_tests.rs:7-180: per-type round trips, with the reference's values -------------------------
def _round_trip(doc):
    data = benc(doc)
    t = OracleOctree.from_bytes(data)
    again = t.to_bytes()
    assert again == data
    return t, bdec(again)


def _pix(color=None, data=None):  # pix_visual / pix_informal / pix_complex (node.rs:354-403)
    return (0xFFFF if color is None else color) | ((0xFFFF if data is None else data) << 16)


def test_node_brickdata_and_nodecontent_serialization():
    """:7-30 (Empty / Solid / Parted bricks) and :32-98 (Nothing, Internal(0xAB), a Leaf with Empty, Solid(pix_complex(69,
    420)) and Parted([pix_visual(666)]) bricks, UniformLeaf(Solid(pix_informal(42)))): decode(encode(x)) == x, checked as
    encode(decode(bytes)) == bytes on a document holding exactly these contents at brick_dim 1."""
    leaf = [None, _pix(69, 420), [_pix(666)], None, None, None, None, None]
    nodes = [(1, None), (1, ("I", 0xAB)), (1, ("L", leaf)), (1, ("U", _pix(None, 42)))]
    doc = tree_doc(True, 4, 1, 4, nodes, [None] * 4, [], [])
    _, back = _round_trip(doc)
    got = [it[1] for it in back[3][1]]
    assert got[0] == b"#" and got[1] == [b"##", 0xAB]
    assert got[2] == [b"###", b"#b", [b"#b#", _pix(69, 420)], [b"##b#", 1, _pix(666), b"#"], b"#b", b"#b", b"#b", b"#b", b"#b"]
    assert got[3] == [b"##u#", [b"#b#", _pix(None, 42)]]
    # a 4x4x4 Parted brick (:11) as a MIP brick of a brick_dim 4 tree
    doc = tree_doc(True, 8, 4, 1, [(1, None)], [None], [(0, 0, 0, 0)], [])
    doc[5] = [["##b#", 64, *([0] * 64), "#"]]
    _, back = _round_trip(doc)
    assert back[5] == [[b"##b#", 64, *([0] * 64), b"#"]]


# :153-180
def test_node_children_serialization():
    doc = tree_doc(True, 4, 1, 3, [(1, None)] * 3, [None, [1, 2, 3, 4, 5, 6, 7, 8], 666], [], [])
    _, back = _round_trip(doc)
    assert back[4] == [b"##x##", [b"##c##", 1, 2, 3, 4, 5, 6, 7, 8], [b"##b##", 666]]


# :100-151
def test_mip_resample_serialization():
    t = OracleOctree(4, 1)
    methods = [(0, 0.0), (1, 0.0), (2, 0.0), (3, 0.420), (4, 0.69)]  # BoxFilter, PointFilter, PointFilterBD, Posterize(0.420), PosterizeBD(0.69)
    for level, (m, thr) in enumerate(methods, start=1):
        t.set_method_at(level, m, thr)
    u = OracleOctree.from_bytes(t.to_bytes())
    for level, (m, thr) in enumerate(methods, start=1):
        assert u.get_method_at(level) == (m, float(np.float32(thr))), level
    # (thr * 1000.) as u32: 0.420 -> code 3 + 420, 0.69 -> 1003 + 690 (bytecode.rs:519-535)
    strategy = bdec(t.to_bytes())[8]
    n = strategy[1]
    codes = dict(zip(strategy[2:2 + 2 * n:2], strategy[3:3 + 2 * n:2]))
    assert codes[4] == 423 and codes[5] == 1693


# ---- src/raytracing/tests.rs:11-66, :90-130: the DDA step against the reference's own plane-intersection form ------
def _plane_line(plane_point, plane_normal, line_origin, line_direction):
    import ctypes as C

    f3 = C.c_float * 3
    d = C.c_float(0)
    some = L.svxo_plane_line_intersection(f3(*plane_point), f3(*plane_normal), f3(*line_origin), f3(*line_direction), C.byref(d))
    return d.value if some else None


def _step_by_planes(cube_min, size, origin, direction):
    """get_step_to_next_sibling (tests.rs:13-66): the far corner of the cube along the ray, the nearest of the three
    planes through it, a step along every axis whose plane is within FLOAT_ERROR_TOLERANCE of the nearest."""
    F = np.float32
    half = F(size) / F(2)
    ref = [F(cube_min[i]) + half + np.copysign(half, F(direction[i])) for i in range(3)]
    f32_max = float(np.finfo(np.float32).max)
    dist = []
    for axis in range(3):
        normal = [1.0 if i == axis else 0.0 for i in range(3)]
        d = _plane_line(ref, normal, origin, direction)
        dist.append(f32_max if d is None else d)
    m = min(dist)
    return [float(np.copysign(1.0, direction[i])) if abs(F(m) - F(dist[i])) < F(1e-5) else 0.0 for i in range(3)]


def test_compare_sibling_step_functions():
    """tests.rs:90-130, `#[ignore = "May fail in edge cases"]` in the reference (it also compares a +-size step with a
    +-1 step, so it can only pass for cubes of size 1). Seeded here, compared by WHICH axes step and in what direction,
    and rays whose two nearest exit planes are closer than 1e-3 are left out: there the two formulations may
    legitimately disagree, which is the reference's stated reason for ignoring the test."""
    import ctypes as C

    f3 = C.c_float * 3
    rng = np.random.default_rng(20240607)
    compared = 0
    for _ in range(400):
        cube_min = [float(rng.integers(0, 100)) for _ in range(3)]
        size = float(rng.integers(1, 1000))
        origin = [float(rng.integers(8, 16)) for _ in range(3)]
        target = np.array(cube_min, dtype=np.float32) + np.float32(size) * np.float32(0.5)
        direction = O.normalized(target - np.array(origin, dtype=np.float32))
        kind, d = C.c_float(), None
        hit = L.svxo_intersect_ray(f3(*cube_min), size, f3(*origin), f3(*direction), C.byref(kind))
        if hit == 0:
            continue
        dist = kind.value if hit == 2 else 0.0  # impact_distance.unwrap_or(0.)
        point = (np.array(origin, dtype=np.float32) + np.array(direction, dtype=np.float32) * np.float32(dist)).astype(np.float32)
        want = _step_by_planes(cube_min, size, origin, direction)
        # distances to the three exit planes from the ray origin: skip near-ties
        F = np.float32
        half = F(size) / F(2)
        ref = [F(cube_min[i]) + half + np.copysign(half, F(direction[i])) for i in range(3)]
        ds = sorted(abs((ref[i] - F(origin[i])) / F(direction[i])) for i in range(3) if direction[i] != 0)
        if len(ds) > 1 and ds[1] - ds[0] < 1e-3 * max(1.0, ds[0]):
            continue
        p = f3(*point)
        step = f3()
        L.svxo_dda_step_to_next_sibling(f3(*origin), f3(*direction), p, f3(*cube_min), size, step)
        assert list(step) == want, (cube_min, size, origin, list(direction))
        # the point moved onto the exit face of the stepped axis
        axis = [i for i in range(3) if want[i] != 0][0]
        assert abs(p[axis] - float(ref[axis])) <= 1e-3 * max(1.0, abs(float(ref[axis])))
        compared += 1
    assert compared > 300
