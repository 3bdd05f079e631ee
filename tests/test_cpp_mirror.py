"""The header-only C++ mirror of the crate API (include/shocovox_b200.hpp) over the C-ABI library: compiled with g++
and run. The construction tests need no GPU; the cpu_render example (the reference's examples/cpu_render.rs on the
B200 path) is a GPU test whose frame must equal the Python-mirror frame byte for byte."""
import subprocess
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest

import shocovox_b200 as S
from shocovox_b200 import scenes

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "examples" / "cpp" / "build"


def compile_example(name: str) -> Path:
    S.lib()
    BUILD.mkdir(exist_ok=True)
    exe = BUILD / name
    src = ROOT / "examples" / "cpp" / f"{name}.cpp"
    lib_dir = S.library_path().parent
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
           f"-L{lib_dir}", "-lshocovox_b200", f"-Wl,-rpath,{lib_dir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def compile_c_example(name: str) -> Path:
    """plain C against include/shocovox_b200.h: the header must be usable without a C++ compiler"""
    S.lib()
    BUILD.mkdir(exist_ok=True)
    exe = BUILD / name
    lib_dir = S.library_path().parent
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "examples" / "cpp" / f"{name}.c"), "-o", str(exe),
           f"-L{lib_dir}", "-lshocovox_b200", f"-Wl,-rpath,{lib_dir}", "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_multi_gpu_example_compiles_as_c():
    compile_c_example("multi_gpu")


@pytest.mark.gpu
def test_c_multi_gpu_two_members():
    """svx_multi_* from plain C with world size 2: two members on device 0, and two real devices when the box has them."""
    exe = compile_c_example("multi_gpu")
    runs = [["0", "0"], ["0", "0", "0", "0"]]
    if S.cuda_device_count() >= 2:
        runs.append(["0", "1"])
    if S.cuda_device_count() >= 8:
        runs.append([str(i) for i in range(8)])
    for devices in runs:
        res = subprocess.run([str(exe)] + devices, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        assert f"multi_gpu: ok ({len(devices)} members" in res.stdout


def test_cpp_construction_api():
    exe = compile_example("octree_api_test")
    vox_path = ROOT / "tests" / "golden" / "vox" / "navigate.vox"
    res = subprocess.run([str(exe), str(vox_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "octree_api_test: ok" in res.stdout
    # Octree::load_vox_file through the C++ mirror builds the tree the Python mirror builds
    want = S.Octree.load_vox_file(str(vox_path), 8)
    line = next(l for l in res.stdout.splitlines() if l.startswith("vox size"))
    assert int(line.split()[2]) == want.get_size() and int(line.split()[4], 16) == want.structure_hash()


def fnv1a(data: bytes, h: int = 1469598103934665603) -> int:
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
def test_cpp_cpu_render_example_matches_python_mirror(tmp_path):
    exe = compile_example("cpu_render")
    out = tmp_path / "frame.ppm"
    cam = scenes.cpu_render_camera()
    cam_args = ["%.9g" % v for v in (*cam.origin, *cam.direction)]  # %.9g round-trips every f32
    res = subprocess.run([str(exe), "150", "150", str(out)] + cam_args, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "detailed_brick_z_edge_error ok" in res.stdout
    line = next(l for l in res.stdout.splitlines() if l.startswith("frame"))
    hits, digest = int(line.split()[3]), int(line.split()[5], 16)
    # the same scene and camera through the Python mirror
    tree = scenes.build_tree(scenes.cpu_render_scene(), S.Octree)
    view = S.OctreeGPUHost(tree).create_new_view(64, S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov), (150, 150))
    f = view.render_to_host()
    assert hits == int((f["hit_id"] != S.MISS).sum())
    want = fnv1a(f["distance"].tobytes(), fnv1a(f["albedo"].tobytes(), fnv1a(f["hit_id"].tobytes())))
    assert digest == want
    assert out.exists() and out.stat().st_size > 150 * 150 * 3


@pytest.mark.gpu
def test_cpp_dot_cube_example_matches_python_mirror(tmp_path):
    """examples/dot_cube.rs through the C++ mirror (tree, viewport, Tab-key CPU render with shading), plain and with MIP
    maps at viewing distance frustum.z; the written image must equal the Python mirror's shaded plane."""
    exe = compile_example("dot_cube")
    light = np.array([0.0, -1.0, 1.0], dtype=np.float32)
    light = (light / np.sqrt((light * light).sum(dtype=np.float32), dtype=np.float32)).astype(np.float32)
    cam = scenes.dot_cube_camera(zoom=True)
    for flag in ([], ["--mips"]):
        out = tmp_path / "dot.ppm"
        res = subprocess.run([str(exe), "96", "64", str(out)] + flag, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout + res.stderr
        tree = S.Octree(256, 32)
        if flag:
            tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
        sc = scenes.dot_cube_scene()
        tree.insert_batch(sc.xyz, sc.rgba)
        view = S.OctreeGPUHost(tree).create_new_view(50, S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov), (96, 64))
        view.set_glass_mode(S.GLASS_AT_FRUSTUM_Z)
        if flag:
            view.set_viewing_distance(200.0)
        view.set_shading(light)
        view.render_to_host()
        want = view.read_shaded()
        rgb = np.stack([want & 0xFF, (want >> 8) & 0xFF, (want >> 16) & 0xFF], axis=-1).astype(np.uint8)
        data = out.read_bytes()
        assert data.endswith(rgb.tobytes()) and len(data) == len(b"P6\n96 64\n255\n") + 96 * 64 * 3
