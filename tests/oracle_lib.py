"""ctypes binding of the CPU ORACLE (oracle/libsvx_oracle.so). Test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module; the product package
(shocovox_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB_PATH = ORACLE_DIR / "libsvx_oracle.so"

EMPTY, VISUAL, INFORMATIVE, COMPLEX = 0, 1, 2, 3
OK, E_INVALID_SIZE, E_INVALID_BRICK_DIMENSION, E_INVALID_STRUCTURE, E_INVALID_POSITION = 0, 1, 2, 3, 4


class Entry(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("rgba", C.c_uint8 * 4), ("data", C.c_uint32)]

    def key(self):
        """Comparable value with the reference's OctreeEntry PartialEq semantics."""
        if self.kind == EMPTY:
            return (EMPTY,)
        if self.kind == VISUAL:
            return (VISUAL, tuple(self.rgba))
        if self.kind == INFORMATIVE:
            return (INFORMATIVE, self.data)
        return (COMPLEX, tuple(self.rgba), self.data)


class Hit(C.Structure):
    _fields_ = [
        ("hit", C.c_uint32),
        ("palette_value", C.c_uint32),
        ("entry", Entry),
        ("impact_point", C.c_float * 3),
        ("normal", C.c_float * 3),
        ("distance", C.c_float),
        ("node_iters", C.c_uint32),
        ("voxel_fetches", C.c_uint32),
        ("outer_iters", C.c_uint32),
        ("would_panic", C.c_uint32),
        ("crawl_iters", C.c_uint32),
        ("mip_probes", C.c_uint32),
    ]


class Camera(C.Structure):
    _fields_ = [
        ("origin", C.c_float * 3),
        ("direction", C.c_float * 3),
        ("glass_width", C.c_float),
        ("glass_height", C.c_float),
        ("glass_distance", C.c_float),
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
        ("node_iters", "<u4"),
        ("voxel_fetches", "<u4"),
        ("outer_iters", "<u4"),
        ("would_panic", "<u4"),
        ("crawl_iters", "<u4"),
        ("mip_probes", "<u4"),
    ]
)
assert HIT_DTYPE.itemsize == C.sizeof(Hit)

ENTRY_DTYPE = np.dtype([("kind", "<u4"), ("rgba", "u1", (4,)), ("data", "<u4")])
assert ENTRY_DTYPE.itemsize == C.sizeof(Entry)


def build(force: bool = False) -> Path:
    srcs = [ORACLE_DIR / n for n in ("svx_oracle.cpp", "svx_oracle_bytecode.cpp", "svx_oracle_capi.cpp", "svx_oracle.hpp", "Makefile")]
    stale = (not LIB_PATH.exists()) or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(LIB_PATH))
    u32, u64, i32, f32, vp = C.c_uint32, C.c_uint64, C.c_int32, C.c_float, C.c_void_p
    f3 = C.POINTER(C.c_float)
    L.svxo_octree_new.argtypes = [u32, u32, C.POINTER(vp)]
    L.svxo_octree_new.restype = i32
    L.svxo_octree_free.argtypes = [vp]
    L.svxo_octree_set_auto_simplify.argtypes = [vp, i32]
    L.svxo_octree_size.argtypes = [vp]
    L.svxo_octree_size.restype = u32
    for name in ("svxo_octree_insert", "svxo_octree_update"):
        getattr(L, name).argtypes = [vp, u32, u32, u32, C.POINTER(Entry)]
        getattr(L, name).restype = i32
    L.svxo_octree_insert_at_lod.argtypes = [vp, u32, u32, u32, u32, C.POINTER(Entry)]
    L.svxo_octree_insert_at_lod.restype = i32
    L.svxo_octree_clear.argtypes = [vp, u32, u32, u32]
    L.svxo_octree_clear.restype = i32
    L.svxo_octree_clear_at_lod.argtypes = [vp, u32, u32, u32, u32]
    L.svxo_octree_clear_at_lod.restype = i32
    L.svxo_octree_insert_batch.argtypes = [vp, vp, vp, vp, u64]
    L.svxo_octree_insert_batch.restype = i32
    L.svxo_octree_get.argtypes = [vp, u32, u32, u32, C.POINTER(Entry)]
    L.svxo_octree_get_sweep.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp]
    L.svxo_octree_to_bytes.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.svxo_bytes_free.argtypes = [vp]
    L.svxo_bytes_free.restype = None
    L.svxo_octree_from_bytes.argtypes = [C.c_char_p, u64, C.POINTER(vp)]
    L.svxo_octree_from_bytes.restype = i32
    L.svxo_octree_structure_hash.argtypes = [vp]
    L.svxo_octree_structure_hash.restype = u64
    L.svxo_octree_node_count.argtypes = [vp]
    L.svxo_octree_node_count.restype = u64
    L.svxo_octree_palette_sizes.argtypes = [vp, C.POINTER(u64)]
    L.svxo_octree_palette_sizes.restype = u64
    L.svxo_octree_get_by_ray.argtypes = [vp, f3, f3, C.POINTER(Hit)]
    L.svxo_octree_get_by_rays.argtypes = [vp, vp, u64, vp]
    L.svxo_octree_get_by_ray_at_lod.argtypes = [vp, f3, f3, f32, C.POINTER(Hit)]
    L.svxo_octree_get_by_rays_at_lod.argtypes = [vp, vp, u64, f32, vp]
    L.svxo_octree_mip_switch.argtypes = [vp, i32]
    L.svxo_octree_mip_enabled.argtypes = [vp]
    L.svxo_octree_mip_enabled.restype = i32
    L.svxo_octree_mip_set_method_at.argtypes = [vp, u64, u32, f32]
    L.svxo_octree_mip_get_method_at.argtypes = [vp, u64, f3]
    L.svxo_octree_mip_get_method_at.restype = u32
    L.svxo_octree_mip_set_color_similarity_thr_at.argtypes = [vp, u64, f32]
    L.svxo_octree_mip_get_color_similarity_at.argtypes = [vp, u64]
    L.svxo_octree_mip_get_color_similarity_at.restype = f32
    L.svxo_octree_mip_reset.argtypes = [vp]
    L.svxo_octree_mip_recalculate.argtypes = [vp]
    L.svxo_octree_mip_sample_root.argtypes = [vp, u32, u32, u32, u32, C.POINTER(Entry)]
    L.svxo_octree_mip_hash.argtypes = [vp]
    L.svxo_octree_mip_hash.restype = u64
    L.svxo_render_rows_lod.argtypes = [vp, C.POINTER(Camera), u32, u32, vp, u32, u32, f32, vp, vp, vp, vp, vp]
    L.svxo_render_rows_lod.restype = C.c_double
    L.svxo_render_rows_shaded.argtypes = [vp, C.POINTER(Camera), u32, u32, vp, u32, u32, f32, vp, vp, vp, vp, vp, vp, vp]
    L.svxo_render_rows_shaded.restype = C.c_double
    L.svxo_make_pixel_ray.argtypes = [C.POINTER(Camera), u32, u32, u32, u32, f3]
    L.svxo_render.argtypes = [vp, C.POINTER(Camera), u32, u32, u32, u32, u32, vp, vp, vp, vp, vp]
    L.svxo_render.restype = C.c_double
    L.svxo_render_rows.argtypes = [vp, C.POINTER(Camera), u32, u32, vp, u32, u32, vp, vp, vp, vp, vp]
    L.svxo_render_rows.restype = C.c_double
    L.svxo_hardware_threads.restype = u32
    L.svxo_hash_region.argtypes = [f32, f32, f32, f32]
    L.svxo_hash_region.restype = u32
    L.svxo_hash_direction.argtypes = [f32, f32, f32]
    L.svxo_hash_direction.restype = u32
    L.svxo_flat_projection.argtypes = [u64] * 4
    L.svxo_flat_projection.restype = u64
    L.svxo_position_in_bitmap_64bits.argtypes = [u64] * 4
    L.svxo_position_in_bitmap_64bits.restype = u64
    L.svxo_set_occupancy_in_bitmap_64bits.argtypes = [u64, u64, u64, u64, u64, i32, u64]
    L.svxo_set_occupancy_in_bitmap_64bits.restype = u64
    L.svxo_child_bounds_for.argtypes = [f3, f32, u32, f3]
    L.svxo_intersect_ray.argtypes = [f3, f32, f3, f3, f3]
    L.svxo_intersect_ray.restype = i32
    L.svxo_plane_line_intersection.argtypes = [f3, f3, f3, f3, f3]
    L.svxo_plane_line_intersection.restype = i32
    L.svxo_cross.argtypes = [f3, f3, f3]
    L.svxo_step_octant.argtypes = [u32, f32, f32, f32]
    L.svxo_step_octant.restype = u32
    L.svxo_cube_impact_normal.argtypes = [f3, f32, f3, f3]
    L.svxo_dda_scale_factors.argtypes = [f3, f3]
    L.svxo_normalized.argtypes = [f3, f3]
    L.svxo_dda_step_to_next_sibling.argtypes = [f3, f3, f3, f3, f32, f3]
    L.svxo_luts.argtypes = [vp, vp, vp, vp, vp]
    L.svxo_node_stack_script.argtypes = [u32, vp, u32, vp]
    L.svxo_node_pool_script.argtypes = [vp, u32, vp]
    L.svxo_brick_is_empty_throughout.argtypes = [u32, u32, vp, u32, i32, u32, vp, u32, vp, u32]
    L.svxo_brick_is_empty_throughout.restype = i32
    _lib = L
    return L


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def albedo_from_u32(value: int):
    """Albedo::from(u32) = 0xRRGGBBAA (src/octree/detail.rs:92-105)."""
    return ((value >> 24) & 0xFF, (value >> 16) & 0xFF, (value >> 8) & 0xFF, value & 0xFF)


def make_entry(albedo=None, data=None) -> Entry:
    """albedo: None | u32 0xRRGGBBAA | (r,g,b,a); data: None | u32."""
    e = Entry()
    if albedo is not None and not isinstance(albedo, tuple):
        albedo = albedo_from_u32(int(albedo))
    if albedo is not None and data is not None:
        e.kind = COMPLEX
    elif albedo is not None:
        e.kind = VISUAL
    elif data is not None:
        e.kind = INFORMATIVE
    else:
        e.kind = EMPTY
    if albedo is not None:
        e.rgba[:] = albedo
    if data is not None:
        e.data = int(data)
    return e


def entry_key(albedo=None, data=None):
    return make_entry(albedo, data).key()


def normalized(v):
    out = (C.c_float * 3)()
    lib().svxo_normalized(_f3(v), out)
    return tuple(out)


class OracleOctree:
    """The reference's `Octree<u32>` API (src/octree/mod.rs, update/insert.rs, raytracing_on_cpu.rs) on the oracle."""

    def __init__(self, size: int, brick_dim: int):
        self._h = C.c_void_p()
        self.status = lib().svxo_octree_new(size, brick_dim, C.byref(self._h))
        if self.status != OK:
            raise ValueError(self.status)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().svxo_octree_free(self._h)
            self._h = C.c_void_p()

    @property
    def handle(self):
        return self._h

    def set_auto_simplify(self, v: bool):
        lib().svxo_octree_set_auto_simplify(self._h, int(v))

    def insert(self, pos, albedo=None, data=None) -> int:
        return lib().svxo_octree_insert(self._h, *pos, C.byref(make_entry(albedo, data)))

    def update(self, pos, albedo=None, data=None) -> int:
        return lib().svxo_octree_update(self._h, *pos, C.byref(make_entry(albedo, data)))

    def insert_at_lod(self, pos, size, albedo=None, data=None) -> int:
        return lib().svxo_octree_insert_at_lod(self._h, *pos, size, C.byref(make_entry(albedo, data)))

    def clear(self, pos) -> int:
        return lib().svxo_octree_clear(self._h, *pos)

    def clear_at_lod(self, pos, size) -> int:
        return lib().svxo_octree_clear_at_lod(self._h, *pos, size)

    def insert_batch(self, xyz: np.ndarray, rgba: np.ndarray, lod: np.ndarray | None = None) -> int:
        xyz = np.ascontiguousarray(xyz, dtype=np.uint32)
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert xyz.shape[1] == 3 and rgba.shape == (xyz.shape[0], 4)
        lod_p = None
        if lod is not None:
            lod = np.ascontiguousarray(lod, dtype=np.uint32)
            lod_p = lod.ctypes.data
        return lib().svxo_octree_insert_batch(self._h, xyz.ctypes.data, rgba.ctypes.data, lod_p, xyz.shape[0])

    def get(self, pos):
        e = Entry()
        lib().svxo_octree_get(self._h, *pos, C.byref(e))
        return e.key()

    def get_sweep(self, origin, extent) -> np.ndarray:
        out = np.zeros(extent[0] * extent[1] * extent[2], dtype=ENTRY_DTYPE)
        lib().svxo_octree_get_sweep(self._h, *origin, *extent, out.ctypes.data)
        return out.reshape(extent)

    def structure_hash(self) -> int:
        return lib().svxo_octree_structure_hash(self._h)

    def to_bytes(self) -> bytes:  # Octree::to_bytes, src/octree/mod.rs:138-142
        buf, n = C.c_void_p(), C.c_uint64()
        lib().svxo_octree_to_bytes(self._h, C.byref(buf), C.byref(n))
        try:
            return C.string_at(buf, n.value)
        finally:
            lib().svxo_bytes_free(buf)

    @classmethod
    def from_bytes(cls, data: bytes) -> "OracleOctree":  # Octree::from_bytes, src/octree/mod.rs:145-148
        h = C.c_void_p()
        status = lib().svxo_octree_from_bytes(bytes(data), len(data), C.byref(h))
        if status != OK:
            raise ValueError(status)
        t = cls.__new__(cls)
        t._h, t.status = h, OK
        return t

    def node_count(self) -> int:
        return lib().svxo_octree_node_count(self._h)

    def get_by_ray(self, origin, direction) -> Hit:
        h = Hit()
        lib().svxo_octree_get_by_ray(self._h, _f3(origin), _f3(direction), C.byref(h))
        return h

    def get_by_rays(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        lib().svxo_octree_get_by_rays(self._h, rays.ctypes.data, rays.shape[0], out.ctypes.data)
        return out

    # ---- MIP maps / LOD (src/octree/mipmap.rs, raytracing_on_cpu.rs:325)
    def get_by_ray_at_lod(self, origin, direction, viewing_distance: float) -> Hit:
        h = Hit()
        lib().svxo_octree_get_by_ray_at_lod(self._h, _f3(origin), _f3(direction), viewing_distance, C.byref(h))
        return h

    def get_by_rays_at_lod(self, rays: np.ndarray, viewing_distance: float) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        lib().svxo_octree_get_by_rays_at_lod(self._h, rays.ctypes.data, rays.shape[0], viewing_distance, out.ctypes.data)
        return out

    def switch_albedo_mip_maps(self, enabled: bool):
        lib().svxo_octree_mip_switch(self._h, 1 if enabled else 0)
        return self

    def mip_enabled(self) -> bool:
        return bool(lib().svxo_octree_mip_enabled(self._h))

    def set_method_at(self, level: int, method: int, thr: float = 0.0):
        lib().svxo_octree_mip_set_method_at(self._h, level, method, thr)
        return self

    def get_method_at(self, level: int):
        thr = C.c_float()
        m = lib().svxo_octree_mip_get_method_at(self._h, level, C.byref(thr))
        return m, thr.value

    def set_color_similarity_thr_at(self, level: int, thr: float):
        lib().svxo_octree_mip_set_color_similarity_thr_at(self._h, level, thr)
        return self

    def get_new_color_similarity_at(self, level: int) -> float:
        return lib().svxo_octree_mip_get_color_similarity_at(self._h, level)

    def mip_reset(self):
        lib().svxo_octree_mip_reset(self._h)
        return self

    def recalculate_mips(self):
        lib().svxo_octree_mip_recalculate(self._h)
        return self

    def sample_root_mip(self, octant: int, pos):
        e = Entry()
        lib().svxo_octree_mip_sample_root(self._h, octant, int(pos[0]), int(pos[1]), int(pos[2]), C.byref(e))
        return e.key()

    def mip_hash(self) -> int:
        return lib().svxo_octree_mip_hash(self._h)

    def render(self, cam: Camera, w: int, h: int, threads: int = 0, rows=None, want_normal=False, row_list=None,
               viewing_distance: float = 3.4028234663852886e38, light_normal=None):
        """Returns dict(hit_id u32[h,w], albedo u8[h,w,4], distance f32[h,w], seconds, counters).
        rows=(r0, r1) renders a contiguous range, row_list an explicit list of image rows (others stay untouched).
        viewing_distance: get_by_ray_at_lod's parameter (default f32::MAX == get_by_ray).
        light_normal: also return "shaded", the RGBA8 pixel of the caller loop (cpu_render.rs:119-136) under that light."""
        r0, r1 = rows if rows is not None else (0, h)
        if row_list is None:
            row_list = np.arange(r0, r1, dtype=np.uint32)
        row_list = np.ascontiguousarray(row_list, dtype=np.uint32)
        hit_id = np.full((h, w), 0xFFFFFFFF, dtype=np.uint32)
        albedo = np.zeros((h, w, 4), dtype=np.uint8)
        dist = np.zeros((h, w), dtype=np.float32)
        normal = np.zeros((h, w, 3), dtype=np.float32) if want_normal else None
        counters = np.zeros(7, dtype=np.uint64)
        shaded = np.full((h, w), 0xFF808080, dtype=np.uint32) if light_normal is not None else None
        light = np.asarray(light_normal if light_normal is not None else (0, 0, 0), dtype=np.float32)
        secs = lib().svxo_render_rows_shaded(
            self._h, C.byref(cam), w, h, row_list.ctypes.data, len(row_list), threads, viewing_distance, hit_id.ctypes.data,
            albedo.ctypes.data, dist.ctypes.data, normal.ctypes.data if want_normal else None, counters.ctypes.data,
            shaded.ctypes.data if shaded is not None else None, light.ctypes.data,
        )
        return {
            "shaded": shaded,
            "hit_id": hit_id, "albedo": albedo, "distance": dist, "normal": normal, "seconds": secs, "rows": row_list,
            "node_iters": int(counters[0]), "voxel_fetches": int(counters[1]), "outer_iters": int(counters[2]),
            "rays_in_root": int(counters[3]), "would_panic": int(counters[4]), "crawl_iters": int(counters[5]),
            "mip_probes": int(counters[6]),
        }


def make_camera(origin, direction, glass_w, glass_h, glass_d) -> Camera:
    c = Camera()
    c.origin[:] = [float(x) for x in origin]
    c.direction[:] = [float(x) for x in direction]
    c.glass_width, c.glass_height, c.glass_distance = float(glass_w), float(glass_h), float(glass_d)
    return c


def pixel_ray(cam: Camera, w, h, x, y) -> np.ndarray:
    out = (C.c_float * 6)()
    lib().svxo_make_pixel_ray(C.byref(cam), w, h, x, y, out)
    return np.array(out[:], dtype=np.float32)


def luts():
    mask = np.zeros(8, dtype=np.uint64)
    index = np.zeros(64, dtype=np.uint32)
    step = np.zeros(27, dtype=np.uint32)
    r2n = np.zeros(512, dtype=np.uint64)
    offs = np.zeros(24, dtype=np.float32)
    lib().svxo_luts(mask.ctypes.data, index.ctypes.data, step.ctypes.data, r2n.ctypes.data, offs.ctypes.data)
    return {"mask": mask, "index": index, "step": step, "ray2node": r2n, "offsets": offs.reshape(8, 3)}
