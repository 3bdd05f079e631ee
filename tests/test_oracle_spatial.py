"""Known-answer tests of the reference's spatial layer, restated against the CPU oracle.

Each test cites the reference test it restates (paths relative to /root/reference/).
"""
import ctypes as C
import json
from pathlib import Path

import numpy as np

import oracle_lib as O

L = O.lib()
GOLDEN = json.loads((Path(__file__).parent / "golden" / "luts.json").read_text())


def f3(*v):
    return (C.c_float * 3)(*v)


# src/raytracing/tests.rs:70-73
def test_special_key_values():
    t = O.OracleOctree(4, 1)
    h = t.get_by_ray((10, 10, 10), O.normalized((1, 1, 1)))
    assert h.hit == 0 and h.palette_value == 4294967295


# src/spatial/tests.rs:23-33
def test_hash_region():
    cases = [((0, 0, 0), 0), ((6, 0, 0), 1), ((0, 0, 6), 2), ((6, 0, 6), 3),
             ((0, 6, 0), 4), ((6, 6, 0), 5), ((0, 6, 6), 6), ((6, 6, 6), 7)]
    for p, want in cases:
        assert L.svxo_hash_region(*[float(v) for v in p], 5.0) == want


# src/spatial/tests.rs:43-70
def test_flat_projection():
    D = 10
    fp = L.svxo_flat_projection
    assert fp(0, 0, 0, D) == 0
    assert fp(10, 0, 0, D) == D
    assert fp(0, 1, 0, D) == D
    assert fp(0, 0, 1, D) == D * D
    assert fp(0, 0, 4, D) == D * D * 4
    assert fp(3, 0, 4, D) == D * D * 4 + 3
    assert fp(3, 2, 4, D) == D * D * 4 + D * 2 + 3
    seen = {fp(x, y, z, D) for x in range(D) for y in range(D) for z in range(D)}
    assert len(seen) == D ** 3


# src/spatial/tests.rs:72-93
def test_position_in_bitmap_64bits():
    p = L.svxo_position_in_bitmap_64bits
    assert p(0, 0, 0, 4) == 0 and p(0, 0, 2, 4) == 32 and p(3, 3, 3, 4) == 63
    assert p(0, 0, 0, 10) == 0 and p(0, 0, 5, 10) == 32 and p(5, 5, 5, 10) == 42 and p(9, 9, 9, 10) == 63
    want = {(0, 0, 0): 0, (1, 0, 0): 2, (0, 1, 0): 8, (1, 1, 0): 10, (0, 0, 1): 32, (1, 0, 1): 34, (0, 1, 1): 40,
            (1, 1, 1): 42}
    for k, v in want.items():
        assert p(*k, 2) == v


# src/spatial/math/tests.rs:89-188
def test_set_occupancy_in_bitmap_64bits():
    s = L.svxo_set_occupancy_in_bitmap_64bits
    m = s(0, 0, 0, 1, 4, 1, 0)
    assert m == 0x0000000000000001
    m = s(3, 3, 3, 1, 4, 1, m)
    assert m == 0x8000000000000001
    m = s(2, 2, 2, 1, 4, 1, m)
    assert m == 0x8000040000000001
    assert s(0, 0, 0, 1, 1, 0, 0xFFFFFFFFFFFFFFFF) == 0
    assert s(0, 0, 0, 1, 1, 1, 0) == 0xFFFFFFFFFFFFFFFF
    m = s(0, 0, 0, 1, 2, 1, 0)
    assert m == 0x0000000000330033
    m = s(1, 1, 1, 1, 2, 1, m)
    assert m == 0xCC00CC0000330033
    assert s(0, 0, 0, 3, 4, 1, 0) == 0x77707770777
    assert s(0, 0, 0, 2, 2, 1, 0) == 0xFFFFFFFFFFFFFFFF
    assert s(0, 0, 0, 5, 4, 1, 0) == 0xFFFFFFFFFFFFFFFF
    assert s(0, 0, 0, 3, 2, 1, 0) == 0xFFFFFFFFFFFFFFFF


# src/spatial/raytracing/tests.rs:41-87
def test_cube_bounds():
    want = [(0, 0, 0), (5, 0, 0), (0, 0, 5), (5, 0, 5), (0, 5, 0), (5, 5, 0), (0, 5, 5), (5, 5, 5)]
    for octant, w in enumerate(want):
        out = (C.c_float * 4)()
        L.svxo_child_bounds_for(f3(0, 0, 0), 10.0, octant, out)
        assert tuple(out[:3]) == tuple(float(v) for v in w) and out[3] == 5.0


def _intersect(min_pos, size, origin, direction):
    d = C.c_float()
    kind = L.svxo_intersect_ray(f3(*min_pos), float(size), f3(*origin), f3(*direction), C.byref(d))
    return kind, d.value


# src/spatial/raytracing/tests.rs:89-198
def test_cube_contains_ray():
    cube = ((0, 0, 0), 4.0)
    assert _intersect(*cube, (2, 5, 2), (0, -1, 0))[0] != 0
    assert _intersect(*cube, (2, -5, 2), (0, 1, 0))[0] != 0
    assert _intersect(*cube, (2, 5, 2), (0, 1, 0))[0] == 0
    assert _intersect(*cube, (-1, -1, -1), O.normalized((1, 1, 1)))[0] != 0
    origin = np.array([4, -1, 4], dtype=np.float32)
    target = np.array([4.055, 4.055, 4.055], dtype=np.float32)
    assert _intersect(*cube, origin, O.normalized(target - origin))[0] == 0
    assert _intersect(*cube, (-1, -1, -1), O.normalized((1, 100, 1)))[0] == 0


# src/spatial/raytracing/tests.rs:200-278
def test_intersect_edge_cases():
    kind, d = _intersect((0, 0, 0), 8.0, (8.0, 4.0, 5.0), (-0.842701, -0.24077171, -0.48154342))
    assert kind == 2 and d == 0.0
    kind, _ = _intersect((0, 0, 0), 16.0, (5.0, 8.0, 5.0), (-0.48507127, -0.7276069, -0.48507127))
    assert kind == 1
    kind, d = _intersect((0, 2, 0), 2.0, (6.0, 7.0, 6.0), (-0.6154574, -0.49236596, -0.6154574))
    assert kind == 2 and d > 0.0


def _plane_line(plane_point, plane_normal, line_origin, line_direction):
    f3 = C.c_float * 3
    d = C.c_float(0)
    some = L.svxo_plane_line_intersection(f3(*plane_point), f3(*plane_normal), f3(*line_origin), f3(*line_direction), C.byref(d))
    return d.value if some else None


# src/spatial/raytracing/tests.rs:6-39
def test_plane_line_intersection():
    assert _plane_line((0, 0, 0), (0, 1, 0), (0, 1, 0), (1, 0, 0)) is None
    assert _plane_line((0, 0, 0), (0, 1, 0), (0, 1, 0), (0, -1, 0)) == 1.0
    assert _plane_line((0, 0, 0), (0, 1, 0), (0, 0, 0), (1, 0, 0)) == 0.0


# src/spatial/math/tests.rs:12-25
def test_negative_intersection():
    assert _plane_line((0, 0, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0)) == -1.0


# src/spatial/tests.rs:7-14
def test_cross_product():
    f3 = C.c_float * 3
    out = f3()
    L.svxo_cross(f3(3, 0, 2), f3(-1, 4, 2), out)
    assert tuple(out) == (-8.0, -8.0, 12.0)


# src/spatial/math/tests.rs:27-51
def test_edge_case_cube_top_hit():
    origin = np.array([8.965594, 10.0, -4.4292345], dtype=np.float32)
    direction = np.array([-0.5082971, -0.72216684, 0.46915793], dtype=np.float32)
    kind, d = _intersect((2.0, 0.0, 0.0), 2.0, origin, direction)
    assert kind == 2 and abs(d - 11.077772) < 0.001
    assert abs((origin + direction * np.float32(d))[1] - 2.0) < 0.001


# src/spatial/math/tests.rs:54-68
def test_impact_normal():
    cases = [((1, 1, 2), (0, 0, 1)), ((1, 2, 1), (0, 1, 0)), ((2, 1, 1), (1, 0, 0)),
             ((1, 1, 0), (0, 0, -1)), ((1, 0, 1), (0, -1, 0)), ((0, 1, 1), (-1, 0, 0))]
    for p, want in cases:
        out = (C.c_float * 3)()
        L.svxo_cube_impact_normal(f3(0, 0, 0), 2.0, f3(*p), out)
        assert tuple(out) == tuple(float(v) for v in want)


# src/spatial/lut.rs:154-896 (literal tables) vs the oracle's regenerated tables
def test_luts_match_reference_literals():
    l = O.luts()
    assert GOLDEN["OOB_OCTANT"] == 8
    assert l["offsets"].tolist() == GOLDEN["OCTANT_OFFSET_REGION_LUT"]
    assert [int(v) for v in l["mask"]] == GOLDEN["BITMAP_MASK_FOR_OCTANT_LUT"]
    assert [int(v) for v in l["index"]] == GOLDEN["BITMAP_INDEX_LUT"]
    assert [int(v) for v in l["step"]] == GOLDEN["OCTANT_STEP_RESULT_LUT"]
    assert [int(v) for v in l["ray2node"]] == GOLDEN["RAY_TO_NODE_OCCUPANCY_BITMASK_LUT"]


# src/spatial/raytracing/mod.rs:68-80 spot checks derived from the LUT definition (lut.rs:91-137)
def test_step_octant():
    assert L.svxo_step_octant(0, 1.0, 0.0, 0.0) == 1
    assert L.svxo_step_octant(1, 1.0, 0.0, 0.0) == 8
    assert L.svxo_step_octant(0, 0.0, 1.0, 0.0) == 4
    assert L.svxo_step_octant(0, 0.0, 0.0, 1.0) == 2
    assert L.svxo_step_octant(7, -1.0, -1.0, -1.0) == 0
    assert L.svxo_step_octant(3, 0.0, 0.0, 0.0) == 3
    assert L.svxo_step_octant(0, -1.0, 0.0, 0.0) == 8


# src/raytracing/tests.rs:817-911 (NodeStack ring buffer)
def _stack(size, ops):
    ops = np.array(ops, dtype=np.int32)
    out = np.zeros(len(ops), dtype=np.int32)
    L.svxo_node_stack_script(size, ops.ctypes.data, len(ops), out.ctypes.data)
    NONE = np.iinfo(np.int32).min
    return [None if v == NONE else int(v) for v in out]


def test_stack_push_and_wrap_around():
    r = _stack(3, [1, 2, 3, 4, -2, -1, -2, -1, -2, -1, -1])
    assert r[4:] == [4, 4, 3, 3, 2, 2, None]


def test_stack_last_and_last_mut():
    assert _stack(3, [10, 20, 30, -2, 40, -2])[3::2] == [30, 40]
    assert _stack(3, [100, 200, 300, -3, -2])[-1] == 350
    assert _stack(4, [-2, -1]) == [None, None]


def test_stack_pop_until_empty():
    assert _stack(3, [5, 15, 25, -1, -1, -1, -1])[3:] == [25, 15, 5, None]
