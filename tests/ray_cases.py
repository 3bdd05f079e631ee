"""The reference's deterministic edge-case rays (src/raytracing/tests.rs:253-813) as data, shared by the oracle
tests (CPU) and the CUDA parity tests (GPU). Each case: a tree recipe, one literal ray, and the reference's assert.

`build(tree)` only uses the Octree API (insert with albedo=/data=), so the same recipe drives the oracle and the product.
"""
import itertools

import numpy as np

F = np.float32


def _norm(v):
    v = np.asarray(v, dtype=F)
    ln = np.sqrt((v[0] * v[0]) + (v[1] * v[1]) + (v[2] * v[2]), dtype=F)
    return v / ln


def _diag_tree(t, c1, c2, c3):
    # tests.rs:255-279 / :319-343 / :489-513 ; Albedo::from(0) is transparent => insert is a no-op
    t.insert((3, 0, 0), 0)
    t.insert((3, 3, 0), 1)
    t.insert((0, 3, 0), 2)
    for y in range(4):
        t.insert((0, y, y), 3)
        t.insert((1, y, y), *c1)
        t.insert((2, y, y), *c2)
        t.insert((3, y, y), *c3)


def _floor_tree(t):
    for x in range(4):
        for z in range(4):
            t.insert((x, 0, z), None, 5)


def _lattice(size, color):
    def build(t):
        q = size // 4
        for x, y, z in itertools.product(range(size), repeat=3):
            if ((x < q or y < q or z < q) and x % 2 == 0 and y % 4 == 0 and z % 2 == 0):
                t.insert((x, y, z), color(x, y, z))
    return build


def _u8(v):
    # Rust `f32 as u8` saturating truncation
    v = F(v)
    return int(min(max(np.trunc(v), 0), 255))


def _grad_color(size):
    def color(x, y, z):
        return (_u8(F(255) * F(x) / F(size)), _u8(F(255) * F(y) / F(size)), _u8(F(255) * F(z) / F(size)), 255)
    return color


def _build_brick_boundary(t):
    S = 128
    q, h = S // 4, S // 2
    for x, y, z in itertools.product(range(S), repeat=3):
        if ((x < q or y < q or z < q) and x % 2 == 0 and y % 4 == 0 and z % 2 == 0) or (h <= x and h <= y and h <= z):
            t.insert((x, y, z), (_u8(F(255) * F(x % 6) / F(6.0)), _u8(F(255) * F(y % 6) / F(6.0)),
                                 _u8(F(255) * F(z % 6) / F(6.0)), 255))


def _build_cube_flaps(t):
    S = 32
    col = _grad_color(S)
    for x, y, z in itertools.product(range(S // 2, S), repeat=3):
        t.insert((x, y, z), col(x, y, z))


def _deep_stack_dir():
    origin = np.array([0.0, 5.0, -1.0], dtype=F)
    target = np.array([511, 511, 511], dtype=F) + F(0.5)
    return origin, _norm(target - origin)


def _behind_dir():
    origin = np.array([2.0, 2.0, -5.0], dtype=F)
    return origin, _norm(np.array([0.0, 3.0, 0.0], dtype=F) - origin)


ALB = lambda v: ("albedo", v)
DATA = lambda v: ("data", v)

# expect: "any" (must only terminate), "hit", "miss", ("albedo", u32), ("data", u32), "miss_or_data5"
CASES = [
    dict(name="unreachable", line=254, size=4, dim=1, build=lambda t: _diag_tree(t, (3,), (3,), (3,)),
         origin=(10.0, 10.0, -5.0), direction=(-0.66739213, -0.6657588, 0.333696), expect="any"),
    dict(name="empty_line_in_middle", line=297, size=4, dim=1, build=lambda t: t.insert((2, 1, 1), 3),
         origin=(8.965594, 10.0, -4.4292345), direction=(-0.5082971, -0.72216684, 0.46915793), expect="hit"),
    dict(name="zero_advance", line=318, size=4, dim=1, build=lambda t: _diag_tree(t, (3,), (3,), (3,)),
         origin=(8.930992, 10.0, -4.498597), direction=(-0.4687217, -0.772969, 0.42757326), expect="hit"),
    dict(name="ray_behind_octree", line=361, size=4, dim=1, build=lambda t: t.insert((0, 3, 0), None, 5),
         origin=_behind_dir()[0], direction=_behind_dir()[1], expect=DATA(5)),
    dict(name="overlapping_voxels", line=378, size=4, dim=1,
         build=lambda t: (t.insert((0, 0, 0), None, 5), t.insert((1, 0, 0), 6)),
         origin=(2.0, 4.0, -2.0), direction=(-0.23184556, -0.79392403, 0.5620785), expect=ALB(6)),
    dict(name="edge_raycast", line=405, size=4, dim=1, build=_floor_tree,
         origin=(2.0, 4.0, -2.0), direction=(-0.47839317, -0.71670955, 0.50741255), expect="miss_or_data5"),
    dict(name="voxel_corner", line=432, size=4, dim=1, build=_floor_tree,
         origin=(2.0, 4.0, -2.0), direction=(-0.27100056, -0.7961219, 0.54106253), expect=DATA(5)),
    dict(name="bottom_edge", line=460, size=4, dim=1, build=_floor_tree,
         origin=(2.0, 4.0, -2.0), direction=(-0.379010856, -0.822795153, 0.423507959), expect=DATA(5)),
    dict(name="loop_stuck", line=488, size=4, dim=1,
         build=lambda t: _diag_tree(t, (4,), (None, 5), (6,)),
         origin=(0.024999974, 10.0, 0.0), direction=(-0.0030831057, -0.98595166, 0.16700225), expect="any"),
    dict(name="brick_undetected", line=531, size=8, dim=4, build=_floor_tree,
         origin=(-1.0716193, 8.0, -7.927902), direction=(0.18699232, -0.6052176, 0.7737865), expect=DATA(5)),
    dict(name="detailed_brick_undetected", line=566, size=8, dim=2,
         build=lambda t: [t.insert(p, None, 5) for p in itertools.product(range(8), repeat=3)],
         origin=(15.8443775, 16.0, 2.226141), direction=(-0.7984906, -0.60134345, 0.028264323), expect=DATA(5)),
    dict(name="detailed_brick_z_edge_error", line=598, size=8, dim=2,
         build=lambda t: [t.insert(p, p[2]) for p in itertools.product(range(1, 8), repeat=3)],
         origin=(11.92238, 16.0, -10.670372), direction=(-0.30062392, -0.6361918, 0.7105529),
         expect=ALB(1), normal=(0.0, 0.0, -1.0)),
    dict(name="deep_stack", line=631, size=512, dim=1,
         build=lambda t: (t.insert((0, 0, 0), 0x000000EE), t.insert((511, 511, 511), 0x000000FF)),
         origin=_deep_stack_dir()[0], direction=_deep_stack_dir()[1], expect=ALB(0x000000FF)),
    dict(name="brick_traversal_error", line=658, size=8, dim=2, build=lambda t: t.insert((0, 0, 0), 0x000000FF),
         origin=(23.84362, 32.0, -21.342018), direction=(-0.51286834, -0.70695364, 0.48701409),
         expect=ALB(0x000000FF), normal_len_lt=1.1),
    dict(name="brick_boundary_error", line=688, size=128, dim=8, build=_build_brick_boundary,
         origin=(191.60886, 256.0, -169.77057), direction=(-0.38838777, -0.49688956, 0.7760514), expect="hit"),
    dict(name="cube_flaps", line=732, size=32, dim=1, build=_build_cube_flaps,
         origin=(47.898006, 64.0, -42.44739), direction=(-0.42279032, -0.4016629, 0.8123516), expect="miss"),
    dict(name="context_bleed", line=773, size=32, dim=1, build=_lattice(32, _grad_color(32)),
         origin=(47.898006, 64.0, -42.44739), direction=(-0.49263135, -0.49703234, 0.714334), expect="hit"),
]


def check_expectation(case, hit, entry_kind, rgba, data, normal):
    """hit: bool; entry_kind 0..3; rgba tuple; data int; normal tuple."""
    exp = case["expect"]
    if exp == "any":
        return
    if exp == "hit":
        assert hit, case["name"]
        return
    if exp == "miss":
        assert not hit, case["name"]
        return
    if exp == "miss_or_data5":
        assert (not hit) or (entry_kind == 2 and data == 5), case["name"]
        return
    kind, value = exp
    assert hit, case["name"]
    if kind == "data":
        assert entry_kind == 2 and data == value, case["name"]
    else:
        want = ((value >> 24) & 0xFF, (value >> 16) & 0xFF, (value >> 8) & 0xFF, value & 0xFF)
        assert entry_kind == 1 and tuple(rgba) == want, case["name"]
    if "normal" in case:
        assert tuple(float(v) for v in normal) == case["normal"], case["name"]
    if "normal_len_lt" in case:
        assert float(np.linalg.norm(np.asarray(normal, dtype=np.float64))) < case["normal_len_lt"]
