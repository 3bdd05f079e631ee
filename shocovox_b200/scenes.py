"""Synthetic scenes and cameras of the BASELINE.json configs, as deterministic voxel lists in insertion order.

Every generator is a pure function of (x, y, z, seed) and returns `(xyz u32[n,3], rgba u8[n,4], lod u32[n] | None)` in the
x -> y -> z ascending order the reference examples insert in (examples/cpu_render.rs:21-43, examples/dot_cube.rs:59-104).
The same arrays feed the product's `Octree.insert_batch` and the test oracle, so both build their tree from the same
insert sequence. Paths are relative to the reference checkout.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

F = np.float32


@dataclass
class Scene:
    name: str
    tree_size: int
    brick_dim: int
    xyz: np.ndarray
    rgba: np.ndarray
    lod: Optional[np.ndarray] = None


@dataclass
class CameraSpec:
    """Viewport fields (src/raytracing/bevy/types.rs:55-71) + which field places the looking glass."""

    origin: Tuple[float, float, float]
    direction: Tuple[float, float, float]
    frustum: Tuple[float, float, float]
    fov: float
    glass_at_frustum_z: bool = False

    @property
    def glass_distance(self) -> float:
        return float(self.frustum[2] if self.glass_at_frustum_z else self.fov)


def _normalized(v) -> np.ndarray:
    v = np.asarray(v, dtype=F)
    ln = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2], dtype=F)
    return (v / ln).astype(F)


def _grid(size_x, size_y=None, size_z=None):
    """All (x, y, z) with x outermost and z innermost, as three flat u32 arrays."""
    size_y = size_x if size_y is None else size_y
    size_z = size_x if size_z is None else size_z
    x, y, z = np.meshgrid(np.arange(size_x, dtype=np.uint32), np.arange(size_y, dtype=np.uint32),
                          np.arange(size_z, dtype=np.uint32), indexing="ij")
    return x.ravel(), y.ravel(), z.ravel()


def _lattice_and_corner(x, y, z, size):
    q, h = size // 4, size // 2
    return (((x < q) | (y < q) | (z < q)) & (x % 2 == 0) & (y % 4 == 0) & (z % 2 == 0)) | ((x >= h) & (y >= h) & (z >= h))


def _f32_to_u8(v: np.ndarray) -> np.ndarray:
    """Rust `f32 as u8`: truncation toward zero, saturating."""
    return np.clip(np.trunc(v), 0, 255).astype(np.uint8)


# ---- C1: examples/cpu_render.rs:13-43 ---------------------------------------------------------------------------
def cpu_render_scene(tree_size: int = 64, brick_dim: int = 8) -> Scene:
    x, y, z = _grid(tree_size)
    m = _lattice_and_corner(x, y, z, tree_size)
    x, y, z = x[m], y[m], z[m]
    ts = F(tree_size)
    rgba = np.stack([_f32_to_u8(F(255) * x.astype(F) / ts), _f32_to_u8(F(255) * y.astype(F) / ts),
                     _f32_to_u8(F(255) * z.astype(F) / ts), np.full(x.shape, 255, np.uint8)], axis=1)
    xyz = np.stack([x, y, z], axis=1)
    # the lone voxel inserted first (cpu_render.rs:10, :22-23): Albedo 0x645097FF at (1, 3, 3)
    xyz = np.concatenate([np.array([[1, 3, 3]], dtype=np.uint32), xyz])
    rgba = np.concatenate([np.array([[0x64, 0x50, 0x97, 0xFF]], dtype=np.uint8), rgba])
    return Scene(f"cpu_render_{tree_size}_{brick_dim}", tree_size, brick_dim, xyz, rgba)


def cpu_render_camera(tree_size: int = 64, k: int = 0) -> CameraSpec:
    """cpu_render.rs:49-94 with the random walk replaced by angle_k = 40 + 0.005 k."""
    radius = F(2.0) * F(tree_size)
    angle = F(40.0) + F(0.005) * F(k)
    origin = np.array([np.sin(angle, dtype=F) * radius, radius, np.cos(angle, dtype=F) * radius], dtype=F)
    direction = _normalized(-origin)
    return CameraSpec(tuple(float(v) for v in origin), tuple(float(v) for v in direction), (4.0, 4.0, 3.0), 3.0)


# ---- C2: examples/dot_cube.rs:24-119 ------------------------------------------------------------------------------
def dot_cube_scene(tree_size: int = 256, brick_dim: int = 32) -> Scene:
    x, y, z = _grid(tree_size)
    m = _lattice_and_corner(x, y, z, tree_size)
    x, y, z = x[m], y[m], z[m]
    q = tree_size // 4
    ts = F(tree_size)

    def channel(c):
        v = np.clip(np.trunc(c.astype(F) / ts * F(255)), 0, 4294967295).astype(np.uint32)  # `as u32`
        return np.where(c % q == 0, v, 128).astype(np.uint8)  # then `as u8` (wraps; values are < 256)

    rgba = np.stack([channel(x), channel(y), channel(z), np.full(x.shape, 255, np.uint8)], axis=1)
    return Scene(f"dot_cube_{tree_size}_{brick_dim}", tree_size, brick_dim, np.stack([x, y, z], axis=1), rgba)


def dot_cube_camera(tree_size: int = 256, zoom: bool = False) -> CameraSpec:
    """dot_cube.rs:48-52, :111-116. `zoom` puts the glass at frustum.z = 200 (dot_cube.rs:209) instead of fov = 3."""
    origin = np.array([tree_size * 2.0, tree_size / 2.0, tree_size * -2.0], dtype=F)
    direction = _normalized(-origin)
    return CameraSpec(tuple(float(v) for v in origin), tuple(float(v) for v in direction), (10.0, 10.0, 200.0), 3.0, zoom)


# ---- benches/performance.rs:11-27 -------------------------------------------------------------------------------------
def criterion_scene(tree_size: int = 512, brick_dim: int = 8, extent: int = 100) -> Scene:
    x, y, z = _grid(extent)
    q, h = tree_size // 4, tree_size // 2
    m = (x < q) | (y < q) | (z < q) | ((x >= h) & (y >= h) & (z >= h))
    x, y, z = x[m], y[m], z[m]
    rgba = np.tile(np.array([[0x00, 0xAB, 0xCD, 0xEF]], dtype=np.uint8), (x.shape[0], 1))
    return Scene(f"criterion_{tree_size}_{brick_dim}", tree_size, brick_dim, np.stack([x, y, z], axis=1), rgba)


# ---- integer hash noise (no library RNG) -------------------------------------------------------------------------------
def _hash2(ix: np.ndarray, iz: np.ndarray, seed: int) -> np.ndarray:
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B1) + iz.astype(np.uint64) * np.uint64(0x85EBCA77) + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    return h


def value_noise(x: np.ndarray, z: np.ndarray, cell: int, seed: int) -> np.ndarray:
    """Bilinear value noise in [0, 1) on a lattice of `cell` voxels; float64 then exact integer thresholds downstream."""
    ix, iz = x // cell, z // cell
    fx, fz = (x % cell) / float(cell), (z % cell) / float(cell)
    sx, sz = fx * fx * (3 - 2 * fx), fz * fz * (3 - 2 * fz)

    def g(a, b):
        return _hash2(a, b, seed).astype(np.float64) / 4294967296.0

    top = g(ix, iz) * (1 - sx) + g(ix + 1, iz) * sx
    bot = g(ix, iz + 1) * (1 - sx) + g(ix + 1, iz + 1) * sx
    return top * (1 - sz) + bot * sz


def terrain_heights(tree_size: int, seed: int, base: int, amplitude: int, block: int = 1) -> np.ndarray:
    """Height (exclusive top y) per (x, z) column: base + floor(amplitude * noise), quantised to `block` voxels."""
    x, z = np.meshgrid(np.arange(tree_size, dtype=np.int64), np.arange(tree_size, dtype=np.int64), indexing="ij")
    if block > 1:
        x, z = (x // block) * block, (z // block) * block
    n = 0.65 * value_noise(x, z, max(tree_size // 16, 4), seed) + 0.35 * value_noise(x, z, max(tree_size // 64, 2), seed + 7)
    h = base + np.floor(amplitude * n).astype(np.int64)
    if block > 1:
        h = (h // block) * block
    return np.clip(h, 1, tree_size).astype(np.uint32)


HEIGHT_BANDS = np.array(
    [[38, 70, 120, 255], [52, 96, 160, 255], [212, 198, 142, 255], [190, 178, 120, 255], [96, 160, 64, 255],
     [80, 144, 56, 255], [64, 128, 48, 255], [52, 110, 44, 255], [110, 100, 80, 255], [128, 116, 96, 255],
     [140, 132, 120, 255], [150, 146, 140, 255], [168, 166, 164, 255], [190, 190, 192, 255], [220, 222, 226, 255],
     [245, 246, 250, 255]], dtype=np.uint8)


def terrain_scene(tree_size: int, brick_dim: int, seed: int, block: int, shell: int = 0, name: str = "terrain") -> Scene:
    """Heightfield terrain (C3: blocky, `block`=4; C5: `block`=1). Columns are filled from y = 0 (or only the top
    `shell` voxels when shell > 0), coloured by height band."""
    base, amplitude = tree_size * 3 // 32, tree_size // 16 * 2 + tree_size // 16
    h = terrain_heights(tree_size, seed, base, amplitude, block)
    n_bands = len(HEIGHT_BANDS)
    xs, ys, zs = [], [], []
    for x in range(tree_size):
        hx = h[x]  # [z]
        top = int(hx.max())
        y, z = np.meshgrid(np.arange(top, dtype=np.uint32), np.arange(tree_size, dtype=np.uint32), indexing="ij")
        m = y < hx[None, :]
        if shell > 0:
            m &= (y + shell) >= hx[None, :]
        ys.append(y[m])
        zs.append(z[m])
        xs.append(np.full(ys[-1].shape, x, dtype=np.uint32))
    x, y, z = np.concatenate(xs), np.concatenate(ys), np.concatenate(zs)
    band = np.minimum((y.astype(np.int64) - base // 2) * n_bands // max(amplitude + base // 2, 1), n_bands - 1)
    band = np.maximum(band, 0)
    rgba = HEIGHT_BANDS[band]
    return Scene(f"{name}_{tree_size}_{brick_dim}", tree_size, brick_dim, np.stack([x, y, z], axis=1), rgba)


def terrain_camera(tree_size: int, pitch_deg: float = 30.0) -> CameraSpec:
    """Above the terrain corner, looking `pitch_deg` down towards the tree centre."""
    origin = np.array([-0.25 * tree_size, 0.75 * tree_size, -0.25 * tree_size], dtype=F)
    target = np.array([0.5 * tree_size, 0.5 * tree_size - np.tan(np.radians(pitch_deg)) * 0.1 * tree_size, 0.5 * tree_size], dtype=F)
    direction = _normalized(target - origin)
    return CameraSpec(tuple(float(v) for v in origin), tuple(float(v) for v in direction), (4.0, 2.25, 3.0), 3.0)


def orbit_cameras(tree_size: int, n: int, radius_factor: float = 1.5, height_factor: float = 0.75):
    """C5: n poses on a circle of radius 1.5 * size at height 0.75 * size, looking at the tree centre."""
    cams = []
    c = np.array([tree_size / 2, tree_size / 2, tree_size / 2], dtype=F)
    for k in range(n):
        a = 2.0 * np.pi * k / n
        origin = np.array([c[0] + radius_factor * tree_size * np.cos(a), height_factor * tree_size,
                           c[2] + radius_factor * tree_size * np.sin(a)], dtype=F)
        direction = _normalized(c - origin)
        cams.append(CameraSpec(tuple(float(v) for v in origin), tuple(float(v) for v in direction), (4.0, 2.25, 3.0), 3.0))
    return cams


# ---- C4: mixed-resolution bricks (insert_at_lod slabs + per-voxel detail) ---------------------------------------------
def colonnade_scene(tree_size: int = 256, brick_dim: int = 8, seed: int = 99) -> Scene:
    """A floor of solid slabs, a colonnade of solid pillars with per-voxel fluting and an arcade of detailed arches.
    Slabs use insert_at_lod sizes of brick_dim .. 4*brick_dim (Solid bricks, UniformLeaf nodes); details are single
    voxels (Parted bricks) - the reference's 'mixed resolution' (README.md:3, src/octree/types.rs:40-52)."""
    S, D = tree_size, brick_dim
    xyz, rgba, lod = [], [], []

    def put(p, c, size=1):
        xyz.append(p)
        rgba.append(c)
        lod.append(size)

    slab = 4 * D
    for x in range(0, S, slab):           # floor: big solid slabs, checkerboard of two stones
        for z in range(0, S, slab):
            c = (176, 168, 150, 255) if ((x // slab) + (z // slab)) % 2 == 0 else (148, 140, 128, 255)
            put((x, 0, z), c, slab)
    pitch = 8 * D
    for px in range(2 * D, S - 2 * D, pitch):   # pillars: stacks of 2D cubes, fluted with single voxels
        for pz in (S // 4, 3 * S // 4):
            for y in range(slab, S // 2, 2 * D):
                put((px, y, pz), (214, 206, 190, 255), 2 * D)
            for y in range(slab, S // 2):
                if _hash2(np.array([px + y]), np.array([pz]), seed)[0] % 3 == 0:
                    put((px - 1, y, pz + (y % (2 * D))), (120, 110, 100, 255))
    for px in range(2 * D, S - 2 * D - pitch, pitch):   # arches between pillars: per-voxel semicircles
        for pz in (S // 4, 3 * S // 4):
            cx, cy, r = px + pitch // 2 + D, S // 2, pitch // 2 - D
            for x in range(px + 2 * D, px + pitch):
                for y in range(cy, cy + r + 2):
                    d2 = (x - cx) ** 2 + (y - cy) ** 2
                    if (r - 2) ** 2 <= d2 <= r ** 2:
                        for dz in range(2 * D):
                            put((x, y, pz + dz), (200, 180 - (dz * 4) % 40, 150, 255))
    order = None  # inserted in generation order (slabs first, then details)
    _ = order
    return Scene(f"colonnade_{S}_{D}", S, D, np.array(xyz, dtype=np.uint32), np.array(rgba, dtype=np.uint8),
                 np.array(lod, dtype=np.uint32))


def colonnade_camera(tree_size: int = 256) -> CameraSpec:
    origin = np.array([-0.2 * tree_size, 0.45 * tree_size, 0.5 * tree_size + 3.0], dtype=F)
    direction = _normalized(np.array([1.0, -0.12, 0.02], dtype=F))
    return CameraSpec(tuple(float(v) for v in origin), tuple(float(v) for v in direction), (4.0, 2.25, 3.0), 3.0)


def build_tree(scene: Scene, tree_cls):
    """Builds `tree_cls(size, brick_dim)` (product Octree or the test oracle) from the scene's insert sequence."""
    t = tree_cls(scene.tree_size, scene.brick_dim)
    t.insert_batch(scene.xyz, scene.rgba, scene.lod)
    return t
