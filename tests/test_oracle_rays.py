"""The reference's ray-query tests (src/raytracing/tests.rs) restated against the CPU oracle."""
import itertools

import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OracleOctree
from ray_cases import CASES, check_expectation

F = np.float32


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_edge_case_rays(case):
    """src/raytracing/tests.rs:253-813 — the crate's only known-answer vectors for get_by_ray."""
    t = OracleOctree(case["size"], case["dim"])
    case["build"](t)
    h = t.get_by_ray(case["origin"], case["direction"])
    assert h.would_panic == 0
    check_expectation(case, bool(h.hit), h.entry.kind, tuple(h.entry.rgba), h.entry.data, tuple(h.normal))


def _ray_to(target, origin):
    d = np.asarray(target, dtype=F) - np.asarray(origin, dtype=F)
    return np.asarray(origin, dtype=F), np.asarray(O.normalized(d), dtype=F)


def _property(size, dim, positions, origin_fn, target_fn, seed):
    rng = np.random.default_rng(seed)
    for trial in range(25):
        t = OracleOctree(size, dim)
        filled = [p for p in positions if rng.integers(0, 20) < 10]
        for p in filled:
            t.insert(p, None, 5)
        for p in filled:
            o, d = _ray_to(target_fn(p), origin_fn(rng))
            h = t.get_by_ray(o, d)
            assert h.hit and h.entry.kind == O.INFORMATIVE and h.entry.data == 5, (trial, p, o, d)
            assert h.would_panic == 0


# src/raytracing/tests.rs:132-142, :145-164 (seeded here; the reference uses thread_rng)
def test_get_by_ray_from_outside():
    pos = [(x, y, 1) for x in range(1, 4) for y in range(1, 4)]
    _property(4, 1, pos, lambda r: tuple(float(v) for v in r.integers(8, 16, 3)), lambda p: p, 1)


# :167-186
def test_get_by_ray_from_outside_where_dim_is_2():
    pos = [(x, y, 1) for x in range(1, 4) for y in range(1, 4)]
    _property(4, 2, pos, lambda r: tuple(float(v) for v in r.integers(8, 16, 3)), lambda p: p, 2)


# :188-225
def test_get_by_ray_from_edge():
    pos = list(itertools.product(range(1, 4), repeat=3))
    _property(8, 1, pos, lambda r: (float(r.integers(0, 8)), float(r.integers(0, 8)), 8.0),
              lambda p: tuple(F(c) + F(0.1) for c in p), 3)


# :227-251
def test_get_by_ray_from_inside():
    pos = list(itertools.product(range(1, 4), repeat=3))
    _property(16, 1, pos, lambda r: tuple(float(v) for v in r.integers(8, 16, 3)), lambda p: p, 4)


def test_deep_stack_uses_restart():
    """the 4-entry ring stack makes the 9-level tree restart from the root."""
    case = next(c for c in CASES if c["name"] == "deep_stack")
    t = OracleOctree(case["size"], case["dim"])
    case["build"](t)
    h = t.get_by_ray(case["origin"], case["direction"])
    assert h.hit and h.outer_iters >= 2


def test_cube_flaps_crawls():
    """a grazing ray is nudged by 0.1 per outer iteration until it leaves the cube."""
    case = next(c for c in CASES if c["name"] == "cube_flaps")
    t = OracleOctree(case["size"], case["dim"])
    case["build"](t)
    h = t.get_by_ray(case["origin"], case["direction"])
    assert not h.hit and h.outer_iters > 50
