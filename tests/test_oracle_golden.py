"""The oracle's outputs on the reference's deterministic cases are frozen in tests/golden/oracle_outputs.json
(tools/make_golden_frames.py): the oracle must keep reproducing them bit for bit, and the CUDA path must match the same
frozen values (GPU test), so the two cannot drift together unnoticed."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from ray_cases import CASES
from shocovox_b200 import scenes

GOLDEN = json.loads((Path(__file__).parent / "golden" / "oracle_outputs.json").read_text())


def bits(v):
    return [int(x) for x in np.asarray(v, dtype=np.float32).view(np.uint32).ravel()]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_frozen_ray_outputs(case):
    g = GOLDEN["rays"][case["name"]]
    t = O.OracleOctree(case["size"], case["dim"])
    case["build"](t)
    assert t.structure_hash() == g["structure_hash"]
    h = t.get_by_ray(case["origin"], case["direction"])
    assert int(h.hit) == g["hit"] and int(h.palette_value) == g["palette_value"]
    assert bits(h.impact_point[:]) == g["impact_point_bits"] and bits([h.distance])[0] == g["distance_bits"]
    if h.hit and not np.isnan(np.asarray(h.normal[:])).any():
        assert bits(h.normal[:]) == g["normal_bits"]
    assert (h.node_iters, h.voxel_fetches, h.outer_iters) == (g["node_iters"], g["voxel_fetches"], g["outer_iters"])


def test_survey_emulation_numbers_are_reproduced():
    """SURVEY.md H2 / H3 (an independent emulation of the reference made during the survey)."""
    deep = GOLDEN["rays"]["deep_stack"]
    assert deep["outer_iters"] == 2
    assert np.array([deep["impact_point_bits"][0]], dtype=np.uint32).view(np.float32)[0] == np.float32(511.00104)
    assert GOLDEN["rays"]["cube_flaps"]["hit"] == 0 and GOLDEN["rays"]["cube_flaps"]["outer_iters"] == 120


def _frame_digests(f):
    return (hashlib.sha256(f["hit_id"].tobytes()).hexdigest(), hashlib.sha256(f["albedo"].tobytes()).hexdigest(),
            hashlib.sha256(f["distance"].tobytes()).hexdigest())


def test_oracle_reproduces_frozen_frames():
    scene = scenes.cpu_render_scene()
    tree = scenes.build_tree(scene, O.OracleOctree)
    assert tree.structure_hash() == GOLDEN["frames"]["scene"]["structure_hash"]
    for k in (0, 21, 63):
        g = GOLDEN["frames"][f"cpu_render_150_k{k}"]
        # the camera comes from the fixture's bit patterns, so the platform's sinf/cosf do not matter
        origin = np.array(g["camera_origin_bits"], dtype=np.uint32).view(np.float32)
        direction = np.array(g["camera_direction_bits"], dtype=np.uint32).view(np.float32)
        f = tree.render(O.make_camera(origin, direction, 4.0, 4.0, 3.0), 150, 150)
        assert _frame_digests(f) == (g["hit_id_sha256"], g["albedo_sha256"], g["distance_sha256"])


@pytest.mark.gpu
def test_gpu_reproduces_frozen_frames():
    import shocovox_b200 as S

    scene = scenes.cpu_render_scene()
    tree = scenes.build_tree(scene, S.Octree)
    assert tree.structure_hash() == GOLDEN["frames"]["scene"]["structure_hash"]
    host = S.OctreeGPUHost(tree)
    for k in (0, 21, 63):
        g = GOLDEN["frames"][f"cpu_render_150_k{k}"]
        # the camera is taken from the fixture's bit patterns, so the platform's libm does not matter
        origin = np.array(g["camera_origin_bits"], dtype=np.uint32).view(np.float32)
        direction = np.array(g["camera_direction_bits"], dtype=np.uint32).view(np.float32)
        view = host.create_new_view(1, S.Viewport(tuple(origin), tuple(direction), (4.0, 4.0, 3.0), 3.0), (150, 150))
        f = view.render_to_host()
        albedo_bytes = f["albedo"].view(np.uint8).reshape(150, 150, 4)
        got = (hashlib.sha256(f["hit_id"].tobytes()).hexdigest(), hashlib.sha256(albedo_bytes.tobytes()).hexdigest(),
               hashlib.sha256(f["distance"].tobytes()).hexdigest())
        assert got == (g["hit_id_sha256"], g["albedo_sha256"], g["distance_sha256"])
