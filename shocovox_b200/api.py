"""Host-side mirror of the reference crate's API for the primary-ray path, over the C ABI (include/shocovox_b200.h).

Names, argument meaning and error behaviour follow shocovox-rs 0.11.1 (paths relative to the reference checkout):

    Octree.new / insert / insert_at_lod / update / get / get_size     src/octree/mod.rs, src/octree/update/insert.rs
    Ray, Octree.get_by_ray                                            src/spatial/raytracing/mod.rs:8-11, src/raytracing/raytracing_on_cpu.rs:316
    Viewport, OctreeGPUHost.create_new_view, OctreeGPUView            src/raytracing/bevy/types.rs:55-130, bevy/data.rs:111

Ray queries run on the GPU only: there is no CPU fallback here. Without a CUDA device `OctreeGPUHost(...)` raises.
torch is optional plumbing (wrapping the device framebuffer for NCCL gathers); the product is the shared library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import build as _build

# ---- status codes (svx_status) ---------------------------------------------------------------------------------
OK = 0
E_INVALID_SIZE = 1
E_INVALID_BRICK_DIMENSION = 2
E_INVALID_STRUCTURE = 3
E_INVALID_POSITION = 4
E_INVALID_ARGUMENT = 5
E_DECODE = 6
E_IO = 7
E_CUDA = -1
E_OUT_OF_MEMORY = -2

ENTRY_EMPTY, ENTRY_VISUAL, ENTRY_INFORMATIVE, ENTRY_COMPLEX = 0, 1, 2, 3
GLASS_AT_FOV, GLASS_AT_FRUSTUM_Z = 0, 1
MISS = 0xFFFFFFFF
F32_MAX = 3.4028234663852886e38  # f32::MAX: the viewing distance of Octree::get_by_ray (raytracing_on_cpu.rs:316-318)
# MIPResamplingMethods (src/octree/types.rs:106-139) as (method, parameter) of the C ABI's svx_mip_method
MIP_BOX_FILTER, MIP_POINT_FILTER, MIP_POINT_FILTER_BD, MIP_POSTERIZE, MIP_POSTERIZE_BD = 0, 1, 2, 3, 4


class OctreeError(Exception):
    """OctreeError of the reference (src/octree/types.rs:9-21) plus device failures."""

    NAMES = {
        E_INVALID_SIZE: "InvalidSize",
        E_INVALID_BRICK_DIMENSION: "InvalidBrickDimension",
        E_INVALID_STRUCTURE: "InvalidStructure",
        E_INVALID_POSITION: "InvalidPosition",
        E_INVALID_ARGUMENT: "InvalidArgument",
        E_DECODE: "Decode",
        E_IO: "Io",
        E_CUDA: "Cuda",
        E_OUT_OF_MEMORY: "OutOfMemory",
    }

    def __init__(self, code: int, message: str = ""):
        self.code = code
        super().__init__(f"{self.NAMES.get(code, code)}: {message}")


class _Albedo(C.Structure):
    _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8), ("a", C.c_uint8)]


class _Entry(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("albedo", _Albedo), ("data", C.c_uint32)]


class _Ray(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("direction", C.c_float * 3)]


class _Hit(C.Structure):
    _fields_ = [
        ("hit", C.c_uint32),
        ("palette_value", C.c_uint32),
        ("entry", _Entry),
        ("impact_point", C.c_float * 3),
        ("normal", C.c_float * 3),
        ("distance", C.c_float),
    ]


class _Viewport(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("direction", C.c_float * 3), ("frustum", C.c_float * 3), ("fov", C.c_float)]


class _Frame(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("row_begin", C.c_uint32),
        ("row_end", C.c_uint32),
        ("hit_id", C.c_void_p),
        ("albedo", C.c_void_p),
        ("distance", C.c_void_p),
        ("kernel_ms", C.c_float),
    ]


class _UploadStats(C.Structure):
    _fields_ = [("bricks", C.c_uint64), ("bytes", C.c_uint64), ("full", C.c_uint32), ("bits_kernel_ms", C.c_float)]


class _GpuStats(C.Structure):
    _fields_ = [
        ("nodes", C.c_uint64),
        ("bricks", C.c_uint64),
        ("voxel_bytes", C.c_uint64),
        ("total_bytes", C.c_uint64),
        ("tree_size", C.c_uint32),
        ("brick_dim", C.c_uint32),
        ("depth", C.c_uint32),
        ("colours", C.c_uint32),
    ]


HIT_DTYPE = np.dtype(
    [
        ("hit", "<u4"),
        ("palette_value", "<u4"),
        ("entry_kind", "<u4"),
        ("rgba", "u1", (4,)),
        ("data", "<u4"),
        ("impact_point", "<f4", (3,)),
        ("normal", "<f4", (3,)),
        ("distance", "<f4"),
    ]
)
ENTRY_DTYPE = np.dtype([("kind", "<u4"), ("rgba", "u1", (4,)), ("data", "<u4")])
VIEWPORT_DTYPE = np.dtype([("origin", "<f4", (3,)), ("direction", "<f4", (3,)), ("frustum", "<f4", (3,)), ("fov", "<f4")])
assert HIT_DTYPE.itemsize == C.sizeof(_Hit) and ENTRY_DTYPE.itemsize == C.sizeof(_Entry)
assert VIEWPORT_DTYPE.itemsize == C.sizeof(_Viewport)

# every symbol include/shocovox_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "svx_version", "svx_last_error_message", "svx_cuda_device_count",
    "svx_octree_new", "svx_octree_free", "svx_octree_insert", "svx_octree_insert_at_lod", "svx_octree_update",
    "svx_octree_clear", "svx_octree_clear_at_lod", "svx_octree_insert_batch", "svx_octree_get", "svx_octree_get_sweep", "svx_octree_size", "svx_octree_brick_dim",
    "svx_octree_set_auto_simplify", "svx_octree_structure_hash", "svx_octree_node_count", "svx_octree_color_palette", "svx_octree_data_palette", "svx_octree_render_data_nodes", "svx_octree_render_data_nodes_with_mips", "svx_octree_render_data_bricks", "svx_render_data_ray_lut",
    "svx_octree_switch_albedo_mip_maps", "svx_octree_mip_maps_enabled", "svx_octree_recalculate_mips",
    "svx_octree_mip_set_method_at", "svx_octree_mip_get_method_at", "svx_octree_mip_set_color_similarity_thr_at",
    "svx_octree_mip_get_color_similarity_at", "svx_octree_mip_reset", "svx_octree_mip_sample_root", "svx_octree_mip_hash",
    "svx_gpu_host_get_by_rays_at_lod", "svx_view_set_viewing_distance", "svx_view_get_viewing_distance",
    "svx_octree_get_by_ray", "svx_octree_get_by_ray_at_lod", "svx_view_reload",
    "svx_view_set_shading", "svx_view_read_shaded", "svx_view_shaded_pointer",
    "svx_octree_to_bytes", "svx_bytes_free", "svx_octree_from_bytes", "svx_octree_save", "svx_octree_load",
    "svx_octree_load_vox", "svx_octree_load_vox_bytes", "svx_vox_required_tree_size", "svx_octree_insert_vox",
    "svx_gpu_host_create", "svx_gpu_host_free", "svx_gpu_host_reload", "svx_gpu_host_last_upload", "svx_gpu_host_stats", "svx_gpu_host_get_by_rays",
    "svx_gpu_host_create_view", "svx_view_free", "svx_view_get_viewport", "svx_view_set_viewport",
    "svx_view_set_glass_mode", "svx_view_set_resolution", "svx_view_resolution", "svx_view_set_shard",
    "svx_view_set_schedule", "svx_view_set_compact_rows", "svx_view_frame_pointers", "svx_view_gather_open", "svx_view_gather_join", "svx_view_gather_join_local",
    "svx_view_gather_close", "svx_view_gather_info",
    "svx_multi_create", "svx_multi_free", "svx_multi_device_count", "svx_multi_view", "svx_multi_set_viewport", "svx_multi_set_glass_mode",
    "svx_multi_set_viewing_distance", "svx_multi_reload", "svx_multi_render", "svx_multi_render_to_host", "svx_multi_render_poses",
    "svx_view_render", "svx_view_read_frame", "svx_view_render_to_host", "svx_view_render_to_host_async", "svx_view_wait_host", "svx_view_render_batch", "svx_view_cuda_stream", "svx_view_device",
    "svx_view_synchronize", "svx_view_timer_start", "svx_view_timer_stop", "svx_view_flush_l2", "svx_view_launch_count",
]

_lib: Optional[C.CDLL] = None


def library_path() -> Path:
    return _build.LIB


def lib() -> C.CDLL:
    """Loads (building first if the sources are newer) the product's shared library. Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    import os

    override = os.environ.get("SVX_LIB")  # a tuning variant built with build.build(out=..., defines=...)
    path = Path(override) if override else _build.build()
    if not path.exists():
        raise OctreeError(E_CUDA, f"{path} is missing: build it with `python -m shocovox_b200.build`")
    L = C.CDLL(str(path))
    u32, u64, i32, vp = C.c_uint32, C.c_uint64, C.c_int32, C.c_void_p
    L.svx_version.restype = C.c_char_p
    L.svx_last_error_message.restype = C.c_char_p
    L.svx_cuda_device_count.restype = i32
    L.svx_octree_new.argtypes = [u32, u32, C.POINTER(vp)]
    L.svx_octree_free.argtypes = [vp]
    L.svx_octree_free.restype = None
    L.svx_octree_insert.argtypes = [vp, u32, u32, u32, C.POINTER(_Entry)]
    L.svx_octree_insert_at_lod.argtypes = [vp, u32, u32, u32, u32, C.POINTER(_Entry)]
    L.svx_octree_update.argtypes = [vp, u32, u32, u32, C.POINTER(_Entry)]
    L.svx_octree_clear.argtypes = [vp, u32, u32, u32]
    L.svx_octree_clear_at_lod.argtypes = [vp, u32, u32, u32, u32]
    L.svx_octree_insert_batch.argtypes = [vp, vp, vp, vp, u64]
    L.svx_octree_get.argtypes = [vp, u32, u32, u32, C.POINTER(_Entry)]
    L.svx_octree_get_sweep.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp]
    L.svx_octree_size.argtypes = [vp]
    L.svx_octree_size.restype = u32
    L.svx_octree_brick_dim.argtypes = [vp]
    L.svx_octree_brick_dim.restype = u32
    L.svx_octree_set_auto_simplify.argtypes = [vp, i32]
    L.svx_octree_to_bytes.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.svx_bytes_free.argtypes = [vp]
    L.svx_bytes_free.restype = None
    L.svx_octree_from_bytes.argtypes = [C.c_char_p, u64, C.POINTER(vp)]
    L.svx_octree_save.argtypes = [vp, C.c_char_p]
    L.svx_octree_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.svx_octree_load_vox.argtypes = [C.c_char_p, u32, C.POINTER(vp)]
    L.svx_octree_load_vox_bytes.argtypes = [C.c_char_p, u64, u32, C.POINTER(vp)]
    L.svx_vox_required_tree_size.argtypes = [C.c_char_p, u64, C.POINTER(u32)]
    L.svx_octree_insert_vox.argtypes = [vp, C.c_char_p, u64]
    L.svx_octree_structure_hash.argtypes = [vp]
    L.svx_octree_structure_hash.restype = u64
    L.svx_octree_node_count.argtypes = [vp]
    L.svx_octree_node_count.restype = u64
    L.svx_octree_color_palette.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.svx_octree_data_palette.argtypes = [vp, vp, u32, C.POINTER(u32)]
    f32 = C.c_float
    L.svx_octree_switch_albedo_mip_maps.argtypes = [vp, i32]
    L.svx_octree_mip_maps_enabled.argtypes = [vp]
    L.svx_octree_recalculate_mips.argtypes = [vp]
    L.svx_octree_mip_set_method_at.argtypes = [vp, u64, i32, f32]
    L.svx_octree_mip_get_method_at.argtypes = [vp, u64, C.POINTER(i32), C.POINTER(f32)]
    L.svx_octree_mip_set_color_similarity_thr_at.argtypes = [vp, u64, f32]
    L.svx_octree_mip_get_color_similarity_at.argtypes = [vp, u64]
    L.svx_octree_mip_get_color_similarity_at.restype = f32
    L.svx_octree_mip_reset.argtypes = [vp]
    L.svx_octree_mip_sample_root.argtypes = [vp, u32, u32, u32, u32, C.POINTER(_Entry)]
    L.svx_octree_mip_hash.argtypes = [vp]
    L.svx_octree_mip_hash.restype = u64
    L.svx_gpu_host_get_by_rays_at_lod.argtypes = [vp, vp, u64, f32, vp]
    L.svx_view_set_viewing_distance.argtypes = [vp, f32]
    L.svx_view_get_viewing_distance.argtypes = [vp, C.POINTER(f32)]
    L.svx_octree_get_by_ray.argtypes = [vp, C.POINTER(_Ray), C.POINTER(_Hit)]
    L.svx_octree_get_by_ray_at_lod.argtypes = [vp, C.POINTER(_Ray), f32, C.POINTER(_Hit)]
    L.svx_view_reload.argtypes = [vp]
    L.svx_view_set_shading.argtypes = [vp, vp]
    L.svx_view_read_shaded.argtypes = [vp, vp]
    L.svx_view_shaded_pointer.argtypes = [vp, C.POINTER(vp)]
    L.svx_gpu_host_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.svx_gpu_host_free.argtypes = [vp]
    L.svx_gpu_host_free.restype = None
    L.svx_gpu_host_reload.argtypes = [vp]
    L.svx_gpu_host_stats.argtypes = [vp, C.POINTER(_GpuStats)]
    L.svx_octree_render_data_nodes.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.svx_octree_render_data_nodes_with_mips.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    L.svx_octree_render_data_bricks.argtypes = [vp, vp, vp, u64, C.POINTER(u64), C.POINTER(u32), C.POINTER(u32)]
    L.svx_render_data_ray_lut.argtypes = [vp]
    L.svx_gpu_host_last_upload.argtypes = [vp, C.POINTER(_UploadStats)]
    L.svx_gpu_host_get_by_rays.argtypes = [vp, vp, u64, vp]
    L.svx_gpu_host_create_view.argtypes = [vp, u32, C.POINTER(_Viewport), u32, u32, C.POINTER(vp)]
    L.svx_view_free.argtypes = [vp]
    L.svx_view_free.restype = None
    L.svx_view_get_viewport.argtypes = [vp, C.POINTER(_Viewport)]
    L.svx_view_set_viewport.argtypes = [vp, C.POINTER(_Viewport)]
    L.svx_view_set_glass_mode.argtypes = [vp, i32]
    L.svx_view_set_resolution.argtypes = [vp, u32, u32]
    L.svx_view_resolution.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.svx_view_set_shard.argtypes = [vp, u32, u32, u32]
    L.svx_view_set_schedule.argtypes = [vp, i32]
    L.svx_view_set_compact_rows.argtypes = [vp, i32]
    L.svx_view_frame_pointers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.svx_view_gather_open.argtypes = [vp, u32, u32, i32, vp]
    L.svx_view_gather_join.argtypes = [vp, u32, vp]
    L.svx_view_gather_join_local.argtypes = [vp, u32, vp]
    L.svx_view_gather_close.argtypes = [vp]
    L.svx_view_gather_info.argtypes = [vp, C.POINTER(i32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.svx_multi_create.argtypes = [vp, C.POINTER(i32), u32, C.POINTER(_Viewport), u32, u32, u32, i32, C.POINTER(vp)]
    L.svx_multi_free.argtypes = [vp]
    L.svx_multi_free.restype = None
    L.svx_multi_device_count.argtypes = [vp]
    L.svx_multi_device_count.restype = u32
    L.svx_multi_view.argtypes = [vp, u32]
    L.svx_multi_view.restype = vp
    L.svx_multi_set_viewport.argtypes = [vp, C.POINTER(_Viewport)]
    L.svx_multi_set_glass_mode.argtypes = [vp, i32]
    L.svx_multi_set_viewing_distance.argtypes = [vp, f32]
    L.svx_multi_reload.argtypes = [vp]
    L.svx_multi_render.argtypes = [vp, C.POINTER(_Frame)]
    L.svx_multi_render_to_host.argtypes = [vp, vp, vp, vp]
    L.svx_multi_render_poses.argtypes = [vp, vp, u32, vp, vp, vp, C.POINTER(C.c_float)]
    L.svx_view_render.argtypes = [vp, C.POINTER(_Frame)]
    L.svx_view_render_to_host.argtypes = [vp, vp, vp, vp]
    L.svx_view_read_frame.argtypes = [vp, vp, vp, vp]
    L.svx_view_render_to_host_async.argtypes = [vp, vp, vp, vp]
    L.svx_view_wait_host.argtypes = [vp, u32, C.POINTER(C.c_float)]
    L.svx_view_render_batch.argtypes = [vp, vp, u32, vp, vp, vp, C.POINTER(C.c_float)]
    L.svx_view_cuda_stream.argtypes = [vp]
    L.svx_view_cuda_stream.restype = vp
    L.svx_view_device.argtypes = [vp]
    L.svx_view_synchronize.argtypes = [vp]
    L.svx_view_timer_start.argtypes = [vp]
    L.svx_view_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.svx_view_flush_l2.argtypes = [vp]
    L.svx_view_launch_count.argtypes = [vp]
    L.svx_view_launch_count.restype = u64
    _lib = L
    return L


WIRE_THREE_PLANES, WIRE_ID_DISTANCE = 0, 1  # svx_wire_format
GATHER_HANDLE_BYTES = 128
E_TIMEOUT = 8


def _check(status: int):
    if status != OK:
        raise OctreeError(status, lib().svx_last_error_message().decode(errors="replace"))


def cuda_device_count() -> int:
    return int(lib().svx_cuda_device_count())


# ---- value types ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Albedo:
    """Albedo{r,g,b,a} (src/octree/types.rs:92-97)."""

    r: int = 0
    g: int = 0
    b: int = 0
    a: int = 0

    @staticmethod
    def from_u32(value: int) -> "Albedo":
        """`Albedo::from(u32)` = 0xRRGGBBAA (src/octree/detail.rs:92-105)."""
        return Albedo((value >> 24) & 0xFF, (value >> 16) & 0xFF, (value >> 8) & 0xFF, value & 0xFF)

    def is_transparent(self) -> bool:
        return self.a == 0


def _as_albedo(v) -> Optional[Albedo]:
    if v is None or isinstance(v, Albedo):
        return v
    if isinstance(v, (tuple, list)):
        return Albedo(*[int(c) for c in v])
    return Albedo.from_u32(int(v))


@dataclass(frozen=True)
class OctreeEntry:
    """OctreeEntry<u32> (src/octree/types.rs:24-36): Empty | Visual(albedo) | Informative(data) | Complex(albedo, data)."""

    albedo: Optional[Albedo] = None
    data: Optional[int] = None

    @property
    def kind(self) -> int:
        if self.albedo is not None and self.data is not None:
            return ENTRY_COMPLEX
        if self.albedo is not None:
            return ENTRY_VISUAL
        if self.data is not None:
            return ENTRY_INFORMATIVE
        return ENTRY_EMPTY

    def is_none(self) -> bool:  # src/octree/mod.rs:102-109
        return (self.albedo is None or self.albedo.is_transparent()) and (self.data is None or self.data == 0)

    def is_some(self) -> bool:
        return not self.is_none()

    def _c(self) -> _Entry:
        e = _Entry()
        e.kind = self.kind
        if self.albedo is not None:
            e.albedo = _Albedo(self.albedo.r, self.albedo.g, self.albedo.b, self.albedo.a)
        if self.data is not None:
            e.data = int(self.data)
        return e

    @staticmethod
    def _from_c(e: _Entry) -> "OctreeEntry":
        alb = Albedo(e.albedo.r, e.albedo.g, e.albedo.b, e.albedo.a) if e.kind in (ENTRY_VISUAL, ENTRY_COMPLEX) else None
        dat = int(e.data) if e.kind in (ENTRY_INFORMATIVE, ENTRY_COMPLEX) else None
        return OctreeEntry(alb, dat)


def entry(albedo=None, data=None) -> OctreeEntry:
    return OctreeEntry(_as_albedo(albedo), None if data is None else int(data))


@dataclass
class Ray:
    """Ray{origin, direction} (src/spatial/raytracing/mod.rs:8-11); direction must be unit length."""

    origin: Sequence[float]
    direction: Sequence[float]


@dataclass
class Viewport:
    """Viewport{origin, direction, frustum, fov} (src/raytracing/bevy/types.rs:55-71)."""

    origin: Sequence[float]
    direction: Sequence[float]
    frustum: Sequence[float] = (4.0, 4.0, 3.0)
    fov: float = 3.0

    def _c(self) -> _Viewport:
        v = _Viewport()
        v.origin[:] = [float(np.float32(x)) for x in self.origin]
        v.direction[:] = [float(np.float32(x)) for x in self.direction]
        v.frustum[:] = [float(np.float32(x)) for x in self.frustum]
        v.fov = float(self.fov)
        return v


@dataclass
class RayHit:
    """The Some((entry, impact_point, normal)) of get_by_ray, plus the palette value and the hit distance."""

    entry: OctreeEntry
    impact_point: tuple
    normal: tuple
    palette_value: int
    distance: float


def normalized(v) -> np.ndarray:
    """V3c::normalized (src/spatial/math/vector.rs:75-81) in f32: v / sqrt((x*x + y*y) + z*z)."""
    v = np.asarray(v, dtype=np.float32)
    ln = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2], dtype=np.float32)
    return (v / ln).astype(np.float32)


# ---- Octree -------------------------------------------------------------------------------------------------------
class Octree:
    """`Octree<u32>`: construction and point queries on the host; ray queries through an OctreeGPUHost."""

    def __init__(self, size: int, brick_dimension: int):
        self._h = C.c_void_p()
        _check(lib().svx_octree_new(size, brick_dimension, C.byref(self._h)))

    @staticmethod
    def new(size: int, brick_dimension: int) -> "Octree":
        return Octree(size, brick_dimension)

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.svx_octree_free(self._h)
            self._h = C.c_void_p()

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def auto_simplify(self):
        raise AttributeError("write-only mirror of the pub field; use set_auto_simplify")

    def set_auto_simplify(self, enabled: bool):
        _check(lib().svx_octree_set_auto_simplify(self._h, int(enabled)))

    def insert(self, position, albedo=None, data=None):
        _check(lib().svx_octree_insert(self._h, *[int(c) for c in position], C.byref(entry(albedo, data)._c())))

    def insert_at_lod(self, position, insert_size: int, albedo=None, data=None):
        _check(lib().svx_octree_insert_at_lod(self._h, *[int(c) for c in position], int(insert_size),
                                              C.byref(entry(albedo, data)._c())))

    def update(self, position, albedo=None, data=None):
        _check(lib().svx_octree_update(self._h, *[int(c) for c in position], C.byref(entry(albedo, data)._c())))

    def clear(self, position):
        _check(lib().svx_octree_clear(self._h, *[int(c) for c in position]))

    def clear_at_lod(self, position, clear_size: int):
        _check(lib().svx_octree_clear_at_lod(self._h, *[int(c) for c in position], int(clear_size)))

    def insert_batch(self, xyz: np.ndarray, rgba: np.ndarray, lod: Optional[np.ndarray] = None):
        xyz = np.ascontiguousarray(xyz, dtype=np.uint32)
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        if xyz.ndim != 2 or xyz.shape[1] != 3 or rgba.shape != (xyz.shape[0], 4):
            raise OctreeError(E_INVALID_ARGUMENT, "xyz must be [n,3] and rgba [n,4]")
        lod_p = None
        if lod is not None:
            lod = np.ascontiguousarray(lod, dtype=np.uint32)
            lod_p = lod.ctypes.data
        _check(lib().svx_octree_insert_batch(self._h, xyz.ctypes.data, rgba.ctypes.data, lod_p, xyz.shape[0]))

    def get(self, position) -> OctreeEntry:
        e = _Entry()
        _check(lib().svx_octree_get(self._h, *[int(c) for c in position], C.byref(e)))
        return OctreeEntry._from_c(e)

    def get_sweep(self, origin, extent) -> np.ndarray:
        out = np.zeros(int(extent[0]) * int(extent[1]) * int(extent[2]), dtype=ENTRY_DTYPE)
        _check(lib().svx_octree_get_sweep(self._h, *[int(c) for c in origin], *[int(c) for c in extent], out.ctypes.data))
        return out.reshape(tuple(int(c) for c in extent))

    def get_size(self) -> int:
        return int(lib().svx_octree_size(self._h))

    def brick_dim(self) -> int:
        return int(lib().svx_octree_brick_dim(self._h))

    def structure_hash(self) -> int:
        return int(lib().svx_octree_structure_hash(self._h))

    def node_count(self) -> int:
        return int(lib().svx_octree_node_count(self._h))

    def color_palette(self) -> np.ndarray:
        """voxel_color_palette (src/octree/types.rs:191) as [n][4] u8 (r, g, b, a): albedo of a hit is color_palette()[hit_id & 0xFFFF]"""
        n = C.c_uint32(0)
        _check(lib().svx_octree_color_palette(self._h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), dtype=np.uint8)
        if n.value:
            _check(lib().svx_octree_color_palette(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def data_palette(self) -> np.ndarray:
        """voxel_data_palette (src/octree/types.rs:192): user data of a hit is data_palette()[hit_id >> 16]"""
        n = C.c_uint32(0)
        _check(lib().svx_octree_data_palette(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        if n.value:
            _check(lib().svx_octree_data_palette(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    # ---- MIP maps: Octree::albedo_mip_map_resampling_strategy() (src/octree/mod.rs:379) ----
    def albedo_mip_map_resampling_strategy(self) -> "StrategyUpdater":
        return StrategyUpdater(self)

    # ---- bencode persistence, src/octree/mod.rs:138-168 ----
    @classmethod
    def _adopt(cls, handle: C.c_void_p) -> "Octree":
        t = cls.__new__(cls)
        t._h = handle
        return t

    def to_bytes(self) -> bytes:
        buf, n = C.c_void_p(), C.c_uint64()
        _check(lib().svx_octree_to_bytes(self._h, C.byref(buf), C.byref(n)))
        try:
            return C.string_at(buf, n.value)
        finally:
            lib().svx_bytes_free(buf)

    @classmethod
    def from_bytes(cls, data: bytes) -> "Octree":
        h = C.c_void_p()
        _check(lib().svx_octree_from_bytes(bytes(data), len(data), C.byref(h)))
        return cls._adopt(h)

    def save(self, path: str):
        _check(lib().svx_octree_save(self._h, str(path).encode()))

    @classmethod
    def load(cls, path: str) -> "Octree":
        h = C.c_void_p()
        _check(lib().svx_octree_load(str(path).encode(), C.byref(h)))
        return cls._adopt(h)

    # `Octree::get_by_ray(&Ray)`: one ray, on the GPU (a host is created on first use and reloaded after edits)
    @classmethod
    def load_vox_file(cls, filename, brick_dimension: int, mip_strategy=None) -> "Octree":
        """`Octree::load_vox_file(filename, brick_dimension)` (src/convert/magicavoxel.rs:266-289); `filename` may also be
        the file's bytes. mip_strategy = callable(StrategyUpdater) applied to the EMPTY tree before the voxels go in -
        `MIPMapStrategy::default().set_enabled(true).load_vox_file(..)` of examples/minecraft.rs:57-60 is
        `mip_strategy=lambda s: s.switch_albedo_mip_maps(True)` (magicavoxel.rs:207-250)."""
        data = bytes(filename) if isinstance(filename, (bytes, bytearray)) else None
        if mip_strategy is None:
            h = C.c_void_p()
            if data is None:
                _check(lib().svx_octree_load_vox(str(filename).encode(), int(brick_dimension), C.byref(h)))
            else:
                _check(lib().svx_octree_load_vox_bytes(data, len(data), int(brick_dimension), C.byref(h)))
            return cls._adopt(h)
        if data is None:
            with open(filename, "rb") as f:
                data = f.read()
        size = C.c_uint32()
        _check(lib().svx_vox_required_tree_size(data, len(data), C.byref(size)))
        tree = cls(int(size.value), int(brick_dimension))
        mip_strategy(tree.albedo_mip_map_resampling_strategy())
        _check(lib().svx_octree_insert_vox(tree.handle, data, len(data)))
        return tree

    def get_by_ray(self, ray: Ray, device: int = 0) -> Optional[RayHit]:
        return self.get_by_ray_at_lod(ray, F32_MAX, device)

    # `Octree::get_by_ray_at_lod(&Ray, viewing_distance)` (src/raytracing/raytracing_on_cpu.rs:325)
    def render_data_nodes(self) -> np.ndarray:
        """Host image of the uploaded node table: [n, 16] u32, one 64-byte record per node in breadth-first order
        (include/shocovox_b200.h: svx_octree_render_data_nodes; the reference's OctreeRenderData, bevy/types.rs:216-279).
        Needs no GPU."""
        n = C.c_uint64(0)
        _check(lib().svx_octree_render_data_nodes(self._h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 16), dtype=np.uint32)
        _check(lib().svx_octree_render_data_nodes(self._h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out

    def render_data_mip_slots(self) -> np.ndarray:
        """Brick slot of every node's MIP brick, in the order of render_data_nodes() (svx_octree_render_data_nodes_with_mips)."""
        n = C.c_uint64(0)
        _check(lib().svx_octree_render_data_nodes_with_mips(self._h, None, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        _check(lib().svx_octree_render_data_nodes_with_mips(self._h, None, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out

    def render_data_bricks(self):
        """Host image of the brick part of the render data: (voxels [n, dim^3] u32, occupancy bits [n, words] u32).
        include/shocovox_b200.h: svx_octree_render_data_bricks. Needs no GPU."""
        n, vol, words = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
        _check(lib().svx_octree_render_data_bricks(self._h, None, None, 0, C.byref(n), C.byref(vol), C.byref(words)))
        voxels = np.zeros((n.value, vol.value), dtype=np.uint32)
        bits = np.zeros((n.value, words.value), dtype=np.uint32)
        _check(lib().svx_octree_render_data_bricks(self._h, voxels.ctypes.data_as(C.c_void_p), bits.ctypes.data_as(C.c_void_p),
                                                   n.value, C.byref(n), C.byref(vol), C.byref(words)))
        return voxels, bits

    def get_by_ray_at_lod(self, ray: Ray, viewing_distance: float, device: int = 0) -> Optional[RayHit]:
        if device == 0:  # svx_octree_get_by_ray_at_lod: the tree handle keeps its own device copy on device 0
            r, h = _Ray(), _Hit()
            r.origin[:] = [float(v) for v in ray.origin]
            r.direction[:] = [float(v) for v in ray.direction]
            _check(lib().svx_octree_get_by_ray_at_lod(self._h, C.byref(r), float(viewing_distance), C.byref(h)))
            if not h.hit:
                return None
            e = OctreeEntry._from_c(h.entry)
            return RayHit(e, tuple(float(v) for v in h.impact_point), tuple(float(v) for v in h.normal),
                          int(h.palette_value), float(h.distance))
        host = getattr(self, "_ray_host", None)
        if host is None or host.device != device:
            host = OctreeGPUHost(self, device)
            self._ray_host = host
        else:
            host.reload()
        return host.get_by_ray(ray, viewing_distance)


class StrategyUpdater:
    """StrategyUpdater<'a, T> (src/octree/types.rs:142, src/octree/mipmap.rs:716-938): chainable MIP map settings of one
    tree, e.g. `tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True).set_method_at(1, MIP_BOX_FILTER)`."""

    def __init__(self, tree: Octree):
        self._tree = tree

    def reset(self) -> "StrategyUpdater":
        _check(lib().svx_octree_mip_reset(self._tree.handle))
        return self

    def is_enabled(self) -> bool:
        return bool(lib().svx_octree_mip_maps_enabled(self._tree.handle))

    def switch_albedo_mip_maps(self, enabled: bool) -> "StrategyUpdater":
        _check(lib().svx_octree_switch_albedo_mip_maps(self._tree.handle, 1 if enabled else 0))
        return self

    def recalculate_mips(self) -> "StrategyUpdater":
        _check(lib().svx_octree_recalculate_mips(self._tree.handle))
        return self

    def get_new_color_similarity_at(self, mip_level: int) -> float:
        return float(lib().svx_octree_mip_get_color_similarity_at(self._tree.handle, int(mip_level)))

    def set_color_similarity_thr_at(self, mip_level: int, similarity_thr: float) -> "StrategyUpdater":
        _check(lib().svx_octree_mip_set_color_similarity_thr_at(self._tree.handle, int(mip_level), float(similarity_thr)))
        return self

    def set_color_similarity_thr(self, levels) -> "StrategyUpdater":
        for mip_level, thr in levels:
            self.set_color_similarity_thr_at(mip_level, thr)
        return self

    def get_method_at(self, mip_level: int):
        """-> (method, parameter); the parameter is the Posterize threshold, 0.0 for the other methods"""
        m, thr = C.c_int32(), C.c_float()
        _check(lib().svx_octree_mip_get_method_at(self._tree.handle, int(mip_level), C.byref(m), C.byref(thr)))
        return int(m.value), float(thr.value)

    def set_method_at(self, mip_level: int, method: int, threshold: float = 0.0) -> "StrategyUpdater":
        _check(lib().svx_octree_mip_set_method_at(self._tree.handle, int(mip_level), int(method), float(threshold)))
        return self

    def set_method(self, levels) -> "StrategyUpdater":
        for mip_level, method in levels:
            if isinstance(method, (tuple, list)):
                self.set_method_at(mip_level, method[0], method[1])
            else:
                self.set_method_at(mip_level, method)
        return self

    def sample_root_mip(self, octant: int, position) -> OctreeEntry:
        """The reference's test hook (mipmap.rs:897-937): MIP voxel of the root (octant 8) or of one of its children."""
        e = _Entry()
        _check(lib().svx_octree_mip_sample_root(self._tree.handle, int(octant), int(position[0]), int(position[1]),
                                                int(position[2]), C.byref(e)))
        return OctreeEntry._from_c(e)

    def mip_hash(self) -> int:
        return int(lib().svx_octree_mip_hash(self._tree.handle))


# ---- OctreeGPUHost / OctreeGPUView ----------------------------------------------------------------------------------
class OctreeGPUHost:
    """OctreeGPUHost{tree} (src/raytracing/bevy/types.rs:80-87): owns the device copy of the whole tree."""

    def __init__(self, tree: Octree, device: int = 0):
        self.tree = tree
        self.device = int(device)
        self._h = C.c_void_p()
        _check(lib().svx_gpu_host_create(tree.handle, self.device, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.svx_gpu_host_free(self._h)
            self._h = C.c_void_p()

    def reload(self):
        _check(lib().svx_gpu_host_reload(self._h))

    def stats(self) -> dict:
        s = _GpuStats()
        _check(lib().svx_gpu_host_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in _GpuStats._fields_}

    def last_upload(self) -> dict:
        """What the most recent upload / reload copied: {"bricks", "bytes", "full", "bits_kernel_ms"}."""
        s = _UploadStats()
        _check(lib().svx_gpu_host_last_upload(self._h, C.byref(s)))
        return {"bricks": int(s.bricks), "bytes": int(s.bytes), "full": bool(s.full), "bits_kernel_ms": float(s.bits_kernel_ms)}

    def get_by_rays(self, rays: np.ndarray, viewing_distance: float = F32_MAX) -> np.ndarray:
        """rays: [n,6] f32 (origin xyz, direction xyz) -> structured array (HIT_DTYPE). viewing_distance is
        get_by_ray_at_lod's parameter; the default (f32::MAX) is get_by_ray."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        _check(lib().svx_gpu_host_get_by_rays_at_lod(self._h, rays.ctypes.data, rays.shape[0], float(viewing_distance),
                                                     out.ctypes.data))
        return out

    def get_by_ray(self, ray: Ray, viewing_distance: float = F32_MAX) -> Optional[RayHit]:
        r = np.concatenate([np.asarray(ray.origin, dtype=np.float32), np.asarray(ray.direction, dtype=np.float32)])
        h = self.get_by_rays(r[None, :], viewing_distance)[0]
        if not h["hit"]:
            return None
        kind = int(h["entry_kind"])
        alb = Albedo(*[int(c) for c in h["rgba"]]) if kind in (ENTRY_VISUAL, ENTRY_COMPLEX) else None
        dat = int(h["data"]) if kind in (ENTRY_INFORMATIVE, ENTRY_COMPLEX) else None
        return RayHit(OctreeEntry(alb, dat), tuple(float(v) for v in h["impact_point"]),
                      tuple(float(v) for v in h["normal"]), int(h["palette_value"]), float(h["distance"]))

    def create_new_view(self, size: int, viewport: Viewport, resolution: Sequence[int]) -> "OctreeGPUView":
        """OctreeGPUHost::create_new_view (src/raytracing/bevy/data.rs:111-166). `size` (node-cache capacity in the
        reference) is accepted and ignored: the whole tree is resident."""
        return OctreeGPUView(self, size, viewport, resolution)


class OctreeGPUView:
    """OctreeGPUView + OctreeSpyGlass (src/raytracing/bevy/types.rs:92-130, bevy/mod.rs:56-99)."""

    def __init__(self, host: OctreeGPUHost, size: int, viewport: Viewport, resolution: Sequence[int]):
        self.host = host
        self._h = C.c_void_p()
        vp = viewport._c()
        _check(lib().svx_gpu_host_create_view(host._h, int(size), C.byref(vp), int(resolution[0]), int(resolution[1]),
                                              C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.svx_view_free(self._h)
            self._h = C.c_void_p()

    def reload(self):
        _check(lib().svx_view_reload(self._h))

    def viewport(self) -> Viewport:
        v = _Viewport()
        _check(lib().svx_view_get_viewport(self._h, C.byref(v)))
        return Viewport(tuple(v.origin), tuple(v.direction), tuple(v.frustum), float(v.fov))

    def set_viewport(self, viewport: Viewport):
        vp = viewport._c()
        _check(lib().svx_view_set_viewport(self._h, C.byref(vp)))

    def set_glass_mode(self, mode: int):
        _check(lib().svx_view_set_glass_mode(self._h, int(mode)))

    def set_shading(self, light_normal=None):
        """Fourth plane = the shaded pixel of the reference's caller loops (examples/cpu_render.rs:119-136) under the
        given diffuse light normal (unit vector; the examples use normalized(0,-1,1)). None switches it off."""
        if light_normal is None:
            _check(lib().svx_view_set_shading(self._h, None))
        else:
            l = np.ascontiguousarray(light_normal, dtype=np.float32)
            _check(lib().svx_view_set_shading(self._h, l.ctypes.data))

    def read_shaded(self) -> np.ndarray:
        """The shaded plane of the last rendered frame, [h, w] RGBA8 packed in u32 (r in the low byte)."""
        w, h = self.resolution()
        out = np.empty((h, w), dtype=np.uint32)
        _check(lib().svx_view_read_shaded(self._h, out.ctypes.data))
        return out

    def set_viewing_distance(self, viewing_distance: float):
        """Viewing distance of every pixel's get_by_ray_at_lod (default f32::MAX = get_by_ray; the reference's GPU
        path uses viewport.frustum.z). Only matters while the tree's MIP maps are enabled."""
        _check(lib().svx_view_set_viewing_distance(self._h, float(viewing_distance)))

    def viewing_distance(self) -> float:
        d = C.c_float()
        _check(lib().svx_view_get_viewing_distance(self._h, C.byref(d)))
        return float(d.value)

    def set_resolution(self, resolution: Sequence[int]):
        _check(lib().svx_view_set_resolution(self._h, int(resolution[0]), int(resolution[1])))

    def resolution(self):
        w, h = C.c_uint32(), C.c_uint32()
        _check(lib().svx_view_resolution(self._h, C.byref(w), C.byref(h)))
        return [int(w.value), int(h.value)]

    def set_shard(self, rank: int, world: int, rows_per_band: int = 8):
        _check(lib().svx_view_set_shard(self._h, int(rank), int(world), int(rows_per_band)))

    def set_schedule(self, persistent):
        """0 / False: static CTAs (default); 1 / True: persistent warps; 2: persistent warps with lane refill (experiment)"""
        _check(lib().svx_view_set_schedule(self._h, int(persistent)))

    def set_compact_rows(self, enabled: bool):
        _check(lib().svx_view_set_compact_rows(self._h, int(enabled)))

    def frame_pointers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().svx_view_frame_pointers(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value or 0), int(b.value or 0), int(c.value or 0)

    # ---- tile-sharded frames, one process per GPU (svx_view_gather_*) ------------------------------------------------
    def gather_open(self, world: int, rows_per_band: int = 8, wire: int = WIRE_THREE_PLANES, export: bool = True) -> Optional[bytes]:
        """Makes this view rank 0 (the assembling GPU) of a `world`-way gather; returns the 128-byte handle to ship to
        the other processes (None with export=False: local peers only)."""
        if not export:
            _check(lib().svx_view_gather_open(self._h, int(world), int(rows_per_band), int(wire), None))
            return None
        buf = (C.c_uint8 * GATHER_HANDLE_BYTES)()
        _check(lib().svx_view_gather_open(self._h, int(world), int(rows_per_band), int(wire), buf))
        return bytes(buf)

    def gather_join(self, rank: int, handle: bytes):
        """Another process' view becomes rank `rank` (1..world-1): its kernel stores into the root's framebuffer."""
        if len(handle) != GATHER_HANDLE_BYTES:
            raise OctreeError(E_INVALID_ARGUMENT, "a gather handle is %d bytes" % GATHER_HANDLE_BYTES)
        buf = (C.c_uint8 * GATHER_HANDLE_BYTES).from_buffer_copy(handle)
        _check(lib().svx_view_gather_join(self._h, int(rank), buf))

    def gather_join_local(self, rank: int, root: "OctreeGPUView"):
        """The same for a root view of this process (peer access or the same device instead of CUDA IPC)."""
        _check(lib().svx_view_gather_join_local(self._h, int(rank), root._h))
        self._gather_root = root  # the root must outlive this membership

    def gather_close(self):
        _check(lib().svx_view_gather_close(self._h))
        self._gather_root = None

    def gather_info(self) -> dict:
        role, rank, world, frames = C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().svx_view_gather_info(self._h, C.byref(role), C.byref(rank), C.byref(world), C.byref(frames)))
        return {"role": ("none", "root", "peer")[role.value], "rank": int(rank.value), "world": int(world.value), "frames": int(frames.value)}

    def render(self, sync: bool = True) -> Optional[dict]:
        """Renders one frame on the device. With sync, returns device pointers and the kernel's CUDA-event time."""
        if not sync:
            _check(lib().svx_view_render(self._h, None))
            return None
        f = _Frame()
        _check(lib().svx_view_render(self._h, C.byref(f)))
        return {"width": f.width, "height": f.height, "hit_id": f.hit_id, "albedo": f.albedo, "distance": f.distance,
                "kernel_ms": float(f.kernel_ms)}

    def render_to_host(self, hit_id: Optional[np.ndarray] = None, albedo: Optional[np.ndarray] = None,
                       distance: Optional[np.ndarray] = None, want=("hit_id", "albedo", "distance")) -> dict:
        """Renders and copies the frame into host arrays (allocated here unless given, e.g. pinned buffers)."""
        w, h = self.resolution()
        if hit_id is None and "hit_id" in want:
            hit_id = np.empty((h, w), dtype=np.uint32)
        if albedo is None and "albedo" in want:
            albedo = np.empty((h, w), dtype=np.uint32)
        if distance is None and "distance" in want:
            distance = np.empty((h, w), dtype=np.float32)
        ptr = lambda a: None if a is None else a.ctypes.data
        _check(lib().svx_view_render_to_host(self._h, ptr(hit_id), ptr(albedo), ptr(distance)))
        return {"hit_id": hit_id, "albedo": albedo, "distance": distance}

    def read_frame(self, want=("hit_id", "albedo", "distance")) -> dict:
        """The framebuffer as it stands (no render): on the root of a gather, the assembled frame of the last render."""
        w, h = self.resolution()
        hit_id = np.empty((h, w), dtype=np.uint32) if "hit_id" in want else None
        albedo = np.empty((h, w), dtype=np.uint32) if "albedo" in want else None
        distance = np.empty((h, w), dtype=np.float32) if "distance" in want else None
        ptr = lambda a: None if a is None else a.ctypes.data
        _check(lib().svx_view_read_frame(self._h, ptr(hit_id), ptr(albedo), ptr(distance)))
        return {"hit_id": hit_id, "albedo": albedo, "distance": distance}

    def render_to_host_ptr(self, hit_id_ptr: int, albedo_ptr: int, distance_ptr: int):
        """Same, into raw host pointers (0 = skip), e.g. torch pinned tensors' data_ptr()."""
        _check(lib().svx_view_render_to_host(self._h, hit_id_ptr or None, albedo_ptr or None, distance_ptr or None))

    def render_to_host_async_ptr(self, hit_id_ptr: int, albedo_ptr: int, distance_ptr: int):
        """Pipelined frame into raw (pinned) host pointers: returns once kernel and copies are queued. The buffers are
        complete only after `wait_host` has retired the frame; at most two frames are in flight."""
        _check(lib().svx_view_render_to_host_async(self._h, hit_id_ptr or None, albedo_ptr or None, distance_ptr or None))

    def wait_host(self, keep_in_flight: int = 0) -> float:
        """Blocks until at most `keep_in_flight` pipelined frames are outstanding; returns (and resets) the summed
        kernel milliseconds of the frames retired since the last call."""
        ms = C.c_float()
        _check(lib().svx_view_wait_host(self._h, keep_in_flight, C.byref(ms)))
        return float(ms.value)

    def render_batch(self, poses: Sequence[Viewport], want=("hit_id", "albedo", "distance")) -> dict:
        w, h = self.resolution()
        n = len(poses)
        arr = np.zeros(n, dtype=VIEWPORT_DTYPE)
        for i, p in enumerate(poses):
            arr[i] = (tuple(np.float32(p.origin)), tuple(np.float32(p.direction)), tuple(np.float32(p.frustum)), p.fov)
        hit_id = np.empty((n, h, w), dtype=np.uint32) if "hit_id" in want else None
        albedo = np.empty((n, h, w), dtype=np.uint32) if "albedo" in want else None
        distance = np.empty((n, h, w), dtype=np.float32) if "distance" in want else None
        ms = C.c_float()
        ptr = lambda a: None if a is None else a.ctypes.data
        _check(lib().svx_view_render_batch(self._h, arr.ctypes.data, n, ptr(hit_id), ptr(albedo), ptr(distance), C.byref(ms)))
        return {"hit_id": hit_id, "albedo": albedo, "distance": distance, "kernel_ms": float(ms.value)}

    def synchronize(self):
        _check(lib().svx_view_synchronize(self._h))

    def timer_start(self):
        _check(lib().svx_view_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _check(lib().svx_view_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        _check(lib().svx_view_flush_l2(self._h))

    def cuda_stream(self) -> int:
        return int(lib().svx_view_cuda_stream(self._h) or 0)

    def launch_count(self) -> int:
        return int(lib().svx_view_launch_count(self._h))


class MultiGPU:
    """svx_multi: one process drives several GPUs - one tree replica and one view per device, ONE frame per render call,
    tile-sharded in row bands and assembled in devices[0]'s framebuffer by the viewport kernels themselves (peer stores +
    device-side flags). The reference has no multi-GPU path; the contract is byte-equality with the single-GPU frame."""

    def __init__(self, tree: Octree, devices: Sequence[int], viewport: Viewport, resolution: Sequence[int], rows_per_band: int = 8,
                 wire: int = WIRE_THREE_PLANES):
        self.tree = tree
        self.devices = [int(d) for d in devices]
        self.resolution = [int(resolution[0]), int(resolution[1])]
        self._h = C.c_void_p()
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        vp = viewport._c()
        _check(lib().svx_multi_create(tree.handle, arr, len(self.devices), C.byref(vp), self.resolution[0], self.resolution[1],
                                      int(rows_per_band), int(wire), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.svx_multi_free(self._h)
            self._h = C.c_void_p()

    def set_viewport(self, viewport: Viewport):
        vp = viewport._c()
        _check(lib().svx_multi_set_viewport(self._h, C.byref(vp)))

    def set_glass_mode(self, mode: int):
        _check(lib().svx_multi_set_glass_mode(self._h, int(mode)))

    def set_viewing_distance(self, viewing_distance: float):
        _check(lib().svx_multi_set_viewing_distance(self._h, float(viewing_distance)))

    def reload(self):
        _check(lib().svx_multi_reload(self._h))

    def render(self, sync: bool = True) -> Optional[dict]:
        if not sync:
            _check(lib().svx_multi_render(self._h, None))
            return None
        f = _Frame()
        _check(lib().svx_multi_render(self._h, C.byref(f)))
        return {"width": f.width, "height": f.height, "hit_id": f.hit_id, "albedo": f.albedo, "distance": f.distance,
                "kernel_ms": float(f.kernel_ms)}

    def read_root_frame(self) -> dict:
        """Renders one frame through the gather and copies the assembled planes from devices[0]."""
        f = self.render(sync=True)
        w, h = self.resolution
        out = {"hit_id": np.empty((h, w), dtype=np.uint32), "albedo": np.empty((h, w), dtype=np.uint32),
               "distance": np.empty((h, w), dtype=np.float32), "kernel_ms": f["kernel_ms"]}
        root = C.c_void_p(lib().svx_multi_view(self._h, 0))
        _check(lib().svx_view_read_frame(root, out["hit_id"].ctypes.data, out["albedo"].ctypes.data, out["distance"].ctypes.data))
        return out

    def render_to_host(self, want=("hit_id", "albedo", "distance")) -> dict:
        w, h = self.resolution
        hit_id = np.empty((h, w), dtype=np.uint32) if "hit_id" in want else None
        albedo = np.empty((h, w), dtype=np.uint32) if "albedo" in want else None
        distance = np.empty((h, w), dtype=np.float32) if "distance" in want else None
        ptr = lambda a: None if a is None else a.ctypes.data
        _check(lib().svx_multi_render_to_host(self._h, ptr(hit_id), ptr(albedo), ptr(distance)))
        return {"hit_id": hit_id, "albedo": albedo, "distance": distance}

    def render_to_host_ptr(self, hit_id_ptr: int, albedo_ptr: int, distance_ptr: int):
        _check(lib().svx_multi_render_to_host(self._h, hit_id_ptr or None, albedo_ptr or None, distance_ptr or None))

    def render_poses(self, poses: Sequence[Viewport], want=("hit_id", "albedo", "distance")) -> dict:
        w, h = self.resolution
        n = len(poses)
        arr = np.zeros(n, dtype=VIEWPORT_DTYPE)
        for i, p in enumerate(poses):
            arr[i] = (tuple(np.float32(p.origin)), tuple(np.float32(p.direction)), tuple(np.float32(p.frustum)), p.fov)
        hit_id = np.empty((n, h, w), dtype=np.uint32) if "hit_id" in want else None
        albedo = np.empty((n, h, w), dtype=np.uint32) if "albedo" in want else None
        distance = np.empty((n, h, w), dtype=np.float32) if "distance" in want else None
        ms = C.c_float()
        ptr = lambda a: None if a is None else a.ctypes.data
        _check(lib().svx_multi_render_poses(self._h, arr.ctypes.data, n, ptr(hit_id), ptr(albedo), ptr(distance), C.byref(ms)))
        return {"hit_id": hit_id, "albedo": albedo, "distance": distance, "kernel_ms": float(ms.value)}
