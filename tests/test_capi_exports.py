"""The C-ABI library loads and exports every symbol include/shocovox_b200.h declares (no GPU needed)."""
import ctypes as C
import re
from pathlib import Path

import pytest

import shocovox_b200 as S
from shocovox_b200 import api

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "shocovox_b200.h").read_text()


def declared_symbols():
    return re.findall(r"SVX_API\s+[\w\s\*]+?\b(svx_\w+)\s*\(", HEADER)


def test_header_declares_what_the_binding_lists():
    assert sorted(declared_symbols()) == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = S.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name


def test_version_and_error_text():
    L = S.lib()
    assert b"sm_100a" in L.svx_version()
    assert isinstance(L.svx_last_error_message(), bytes)


def test_octree_errors_follow_the_reference_order():
    # src/octree/mod.rs:174-187
    for size, dim, code in [(0, 8, api.E_INVALID_BRICK_DIMENSION), (64, 3, api.E_INVALID_BRICK_DIMENSION),
                            (4, 8, api.E_INVALID_SIZE), (24, 8, api.E_INVALID_SIZE), (8, 8, api.E_INVALID_STRUCTURE)]:
        with pytest.raises(S.OctreeError) as e:
            S.Octree(size, dim)
        assert e.value.code == code
    t = S.Octree(4, 1)
    with pytest.raises(S.OctreeError) as e:
        t.insert((4, 0, 0), 0xFF0000FF)
    assert e.value.code == api.E_INVALID_POSITION
    assert t.get_size() == 4 and t.brick_dim() == 1


def test_null_arguments_are_rejected_not_crashed():
    L = S.lib()
    assert L.svx_octree_new(4, 1, None) == api.E_INVALID_ARGUMENT
    assert L.svx_gpu_host_create(None, 0, None) == api.E_INVALID_ARGUMENT
    assert L.svx_view_render(None, None) == api.E_INVALID_ARGUMENT


def test_multi_gpu_and_vox_entry_points_check_their_arguments():
    """The gather / multi / .vox entry points refuse null handles and malformed input with a status, on a box without a GPU too."""
    import ctypes as C

    L = S.lib()
    for call in (lambda: L.svx_view_gather_open(None, 2, 8, 0, None), lambda: L.svx_view_gather_join(None, 1, None),
                 lambda: L.svx_view_gather_join_local(None, 1, None), lambda: L.svx_view_gather_close(None),
                 lambda: L.svx_view_gather_info(None, None, None, None, None), lambda: L.svx_view_read_frame(None, None, None, None),
                 lambda: L.svx_multi_create(None, None, 2, None, 64, 64, 8, 0, None), lambda: L.svx_multi_render(None, None),
                 lambda: L.svx_multi_render_to_host(None, None, None, None), lambda: L.svx_multi_set_viewport(None, None),
                 lambda: L.svx_multi_render_poses(None, None, 1, None, None, None, None),
                 lambda: L.svx_octree_load_vox(None, 8, None), lambda: L.svx_octree_load_vox_bytes(None, 0, 8, None),
                 lambda: L.svx_vox_required_tree_size(None, 0, None), lambda: L.svx_octree_insert_vox(None, None, 0)):
        assert call() == api.E_INVALID_ARGUMENT
    L.svx_multi_free(None)  # a null handle is a no-op, like the other *_free
    assert L.svx_multi_device_count(None) == 0 and L.svx_multi_view(None, 0) is None
    out = C.c_void_p()
    assert L.svx_octree_load_vox_bytes(b"VOX \x96\x00\x00\x00", 8, 8, C.byref(out)) == api.E_DECODE and not out.value
    t = S.Octree(64, 8)
    # svx_multi_create validates the shape before it touches a device
    vp = S.Viewport((100.0, 100.0, 100.0), tuple(S.normalized((-1, -1, -1))))._c()
    devs = (C.c_int32 * 2)(0, 0)
    for world, band, wire in ((0, 8, 0), (17, 8, 0), (2, 6, 0), (2, 8, 7)):
        assert L.svx_multi_create(t.handle, devs, world, C.byref(vp), 64, 64, band, wire, C.byref(out)) == api.E_INVALID_ARGUMENT
    # get_sweep: a box that leaves the tree is refused instead of wrapping around in 32 bits
    buf = (C.c_uint8 * 64)()
    assert L.svx_octree_get_sweep(t.handle, 60, 0, 0, 8, 1, 1, buf) == api.E_INVALID_POSITION
    assert L.svx_octree_get_sweep(t.handle, 0xFFFFFFFF, 0, 0, 2, 1, 1, buf) == api.E_INVALID_POSITION


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the ray path must fail loudly, never fall back to host code."""
    if S.cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    t = S.Octree(4, 1)
    t.insert((1, 1, 1), 0xFF0000FF)
    with pytest.raises(S.OctreeError) as e:
        S.OctreeGPUHost(t)
    assert e.value.code == api.E_CUDA
    with pytest.raises(S.OctreeError):
        t.get_by_ray(S.Ray((5, 5, 5), tuple(S.normalized((-1, -1, -1)))))


def test_product_does_not_reference_the_oracle():
    """The shipped package must not import, link or call anything under oracle/."""
    pkg = ROOT / "shocovox_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.hpp")) + list(pkg.rglob("*.cuh")):
        text = p.read_text()
        assert "svxo_" not in text and "oracle_lib" not in text and "libsvx_oracle" not in text, p


def test_built_library_has_no_packed_fused_multiply_add():
    """The DDA uses packed f32 add / mul (FADD2 / FMUL2); a packed FMA (FFMA2) would round once where the reference
    (raytracing_on_cpu.rs:124-152, Rust never contracts) rounds twice. build.py refuses such a library; check the one in the tree."""
    import shutil
    import subprocess

    from shocovox_b200 import build

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([tool, "-sass", str(build.LIB)], capture_output=True, text=True, check=True).stdout
    assert "FADD2" in sass and "FMUL2" in sass, "the packed f32 path is not in the built kernels"
    assert "FFMA2" not in sass
    build.check_no_packed_fma(build.LIB)


def test_packed_fma_check_refuses_a_contracted_library(tmp_path):
    """What the check exists for: ptxas 12.9 contracts mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 although every operand says
    .rn and --fmad=false is given. A library built from exactly that pattern must be refused (and removed)."""
    import subprocess

    from shocovox_b200 import build

    src = tmp_path / "contract.cu"
    src.write_text(
        '#include <cstdint>\n'
        'extern "C" __global__ void k(const uint64_t* a, const uint64_t* b, uint64_t* c) {\n'
        '    uint64_t p, s;\n'
        '    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(a[threadIdx.x]), "l"(b[threadIdx.x]));\n'
        '    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(p), "l"(c[threadIdx.x]));\n'
        '    c[threadIdx.x] = s;\n'
        '}\n')
    lib = tmp_path / "libcontract.so"
    subprocess.run([build.nvcc_path(), "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
                    "-cudart", "static", "-o", str(lib), str(src)], check=True, capture_output=True)
    with pytest.raises(RuntimeError, match="packed fused multiply-add"):
        build.check_no_packed_fma(lib)
    assert not lib.exists()
