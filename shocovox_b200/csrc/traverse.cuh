// Device-side primary-ray traversal of the brick-leaf sparse voxel octree: Octree::get_by_ray of the reference
// (src/raytracing/raytracing_on_cpu.rs:316-565) restated for sm_100a. NOT a port of the WGSL shader (which is a
// different, lossy variant - SURVEY §2.2); results must equal the CPU get_by_ray bit for bit.
//
// Numerics: compiled with -fmad=false (Rust never contracts to FMA), IEEE division / sqrt (nvcc defaults
// -prec-div=true -prec-sqrt=true, -ftz=false). Divisions by node / brick sizes are by exact powers of two and are
// written as multiplications by the exact reciprocal: x / 2^k and x * 2^-k are the same real number, both correctly
// rounded, hence the same bits. Node bounds are exact small integers in f32.
//
// The reference's look-up tables (src/spatial/lut.rs) are replaced by closed forms of their generator logic
// (lut.rs:12-152): no table loads on the hot path. gpu_selftest.cu checks every table entry against the closed forms.
// `file:line` citations are relative to the reference checkout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gpu_tree.hpp"

namespace svx {

constexpr uint32_t OOB_OCTANT = 8;  // lut.rs:154
constexpr float FLOAT_ERROR_TOLERANCE = 0.00001f;  // spatial/raytracing/mod.rs:5

struct TraceResult {
    uint32_t palette_value;  // NIL on a miss
    float px, py, pz;        // impact point
    float bx, by, bz, bsize; // bounds of the hit cell (for the normal)
};

// f32::clamp(min, max): NaN stays NaN
__device__ __forceinline__ float rust_clamp(float v, float lo, float hi) {
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

// f32::signum: +-1 with the sign bit of v (also for +-0), NaN for NaN
__device__ __forceinline__ float rust_signum(float v) { return (v != v) ? v : copysignf(1.0f, v); }

// hash_region, spatial/math/mod.rs:11-19
__device__ __forceinline__ uint32_t hash_region(float x, float y, float z, float half) {
    return (uint32_t)(x >= half) + ((uint32_t)(z >= half) << 1) + ((uint32_t)(y >= half) << 2);
}

// BITMAP_MASK_FOR_OCTANT_LUT[o] (lut.rs:199-208) = 0x0000000000330033 << (2*xbit + 8*ybit + 32*zbit)
__device__ __forceinline__ bool octant_occupied(uint32_t oc_lo, uint32_t oc_hi, uint32_t octant) {
    const uint32_t half = (octant & 2u) ? oc_hi : oc_lo;  // z bit selects the upper 32 cells
    const uint32_t m = 0x00330033u << (((octant & 1u) << 1) + ((octant & 4u) << 1));
    return (half & m) != 0u;
}

// RAY_TO_NODE_OCCUPANCY_BITMASK_LUT[cell][dir] (generate_lut_64_bits, lut.rs:39-89): all cells of the 4x4x4 bitmap
// inside the box spanned from `cell` to the corner the direction octant points at. cell = x + 4y + 16z
// (BITMAP_INDEX_LUT, lut.rs:210-235); dirbits: bit0 x+, bit1 z+, bit2 y+ (hash_direction, math/mod.rs:22-26).
__device__ __forceinline__ bool ray_may_hit(uint32_t oc_lo, uint32_t oc_hi, uint32_t cx, uint32_t cy, uint32_t cz,
                                            uint32_t dirbits) {
    const uint32_t xn = (dirbits & 1u) ? ((0xFu << cx) & 0xFu) : (0xFu >> (3u - cx));
    const uint32_t yr = (dirbits & 4u) ? ((0xFFFFu << (4u * cy)) & 0xFFFFu) : (0xFFFFu >> (4u * (3u - cy)));
    const uint32_t plane = (xn * 0x1111u) & yr;   // 16 cells of one z-slab
    const uint32_t both = plane * 0x00010001u;     // the same in both slabs of a 32-bit half
    // z-slabs 0,1 live in oc_lo, 2,3 in oc_hi
    const uint32_t zsel = (dirbits & 2u) ? ((0xFu << cz) & 0xFu) : (0xFu >> (3u - cz));
    const uint32_t mlo = ((zsel & 1u) ? 0x0000FFFFu : 0u) | ((zsel & 2u) ? 0xFFFF0000u : 0u);
    const uint32_t mhi = ((zsel & 4u) ? 0x0000FFFFu : 0u) | ((zsel & 8u) ? 0xFFFF0000u : 0u);
    return ((oc_lo & both & mlo) | (oc_hi & both & mhi)) != 0u;
}

// step_octant, spatial/raytracing/mod.rs:68-80 with OCTANT_STEP_RESULT_LUT (generate_octant_step_result_lut,
// lut.rs:91-137): move one octant along each stepped axis, OOB when leaving the 2x2x2 block.
// `step` components are `(v as i32).signum()` of a +-1.0 / 0.0 / NaN float.
__device__ __forceinline__ uint32_t step_octant(uint32_t octant, float sx, float sy, float sz) {
    const int ix = (int)(octant & 1u) + (sx > 0.5f) - (sx < -0.5f);
    const int iz = (int)((octant >> 1) & 1u) + (sz > 0.5f) - (sz < -0.5f);
    const int iy = (int)((octant >> 2) & 1u) + (sy > 0.5f) - (sy < -0.5f);
    if (((ix | iy | iz) & ~1) != 0) return OOB_OCTANT;
    return (uint32_t)(ix | (iz << 1) | (iy << 2));
}

struct RayConst {
    float ox, oy, oz;      // origin
    float dx, dy, dz;      // direction
    float sfx, sfy, sfz;   // get_dda_scale_factors, raytracing_on_cpu.rs:99-112
    float sgx, sgy, sgz;   // signum(direction)
    float s0x, s0y, s0z;   // signum.max(0.)
    uint32_t dirbits;      // hash_direction
};

// dda_step_to_next_sibling, raytracing_on_cpu.rs:124-152
__device__ __forceinline__ void dda_step(const RayConst& r, float& px, float& py, float& pz, float bx, float by,
                                         float bz, float bsize, float& stx, float& sty, float& stz) {
    const float dfx = px - bx, dfy = py - by, dfz = pz - bz;
    const float nx = bsize * r.s0x - r.sgx * dfx;
    const float ny = bsize * r.s0y - r.sgy * dfy;
    const float nz = bsize * r.s0z - r.sgz * dfz;
    const float d_x = fabsf(nx * r.sfx);
    const float d_y = fabsf(ny * r.sfy);
    const float d_z = fabsf(nz * r.sfz);
    const float m = fminf(fminf(d_x, d_y), d_z);
    px = px + r.dx * m;
    py = py + r.dy * m;
    pz = pz + r.dz * m;
    stx = (m == d_x) ? r.sgx : 0.0f;
    sty = (m == d_y) ? r.sgy : 0.0f;
    stz = (m == d_z) ? r.sgz : 0.0f;
}

__device__ __forceinline__ void ray_setup(RayConst& r) {
    auto sq = [](float v) { return v * v; };  // `.powf(2.)` == x*x
    r.sfx = sqrtf(1.0f + sq(r.dz / r.dx) + sq(r.dy / r.dx));
    r.sfy = sqrtf(sq(r.dx / r.dy) + 1.0f + sq(r.dz / r.dy));
    r.sfz = sqrtf((sq(r.dx / r.dz) + 1.0f) + sq(r.dy / r.dz));
    r.sgx = rust_signum(r.dx);
    r.sgy = rust_signum(r.dy);
    r.sgz = rust_signum(r.dz);
    r.s0x = fmaxf(r.sgx, 0.0f);
    r.s0y = fmaxf(r.sgy, 0.0f);
    r.s0z = fmaxf(r.sgz, 0.0f);
    r.dirbits = hash_region(1.0f + r.dx, 1.0f + r.dy, 1.0f + r.dz, 1.0f);
}

// `(v as i32).clamp(0, dim-1)`: cvt.rzi saturates and maps NaN to 0 like Rust's `as`
__device__ __forceinline__ int clamp_index(float v, int dim) { return min(max(__float2int_rz(v), 0), dim - 1); }
// `v.floor() as usize` for the 4x4x4 bitmap position; the reference bounds-panics above 3, we clamp
__device__ __forceinline__ uint32_t bitmap_coord(float v) { return (uint32_t)min(max(__float2int_rd(v), 0), 3); }

// traverse_brick, raytracing_on_cpu.rs:156-252. Walks the occupancy bit-brick; returns the flat index of the first
// non-empty voxel or -1.
__device__ __forceinline__ int traverse_brick(const DeviceTree& t, const RayConst& r, float& px, float& py, float& pz,
                                              uint32_t brick, float bx, float by, float bz, float bsize,
                                              float inv_size, int& hx, int& hy, int& hz) {
    const int dim = (int)t.brick_dim;
    const float fdim = (float)dim;
    int ix = clamp_index((px - bx) * fdim * inv_size, dim);
    int iy = clamp_index((py - by) * fdim * inv_size, dim);
    int iz = clamp_index((pz - bz) * fdim * inv_size, dim);
    const float unit = bsize * t.inv_brick_dim;  // size / dim, exact: both powers of two
    float cx = bx + (float)ix * unit, cy = by + (float)iy * unit, cz = bz + (float)iz * unit;
    const uint32_t* bits = t.brick_bits + (size_t)brick * t.bit_words;
    const uint32_t sh = t.brick_shift;
    for (;;) {
        if (((ix | iy | iz) < 0) || ix >= dim || iy >= dim || iz >= dim) return -1;
        const int flat = ix + (iy << sh) + (iz << (2 * sh));
        if ((__ldg(bits + (flat >> 5)) >> (flat & 31)) & 1u) {
            hx = ix; hy = iy; hz = iz;
            return flat;
        }
        float stx, sty, stz;
        dda_step(r, px, py, pz, cx, cy, cz, unit, stx, sty, stz);
        cx = cx + stx * unit;
        cy = cy + sty * unit;
        cz = cz + stz * unit;
        ix += __float2int_rn(stx);  // V3c::<i32>::from(step) rounds (vector.rs:354-364); NaN -> 0
        iy += __float2int_rn(sty);
        iz += __float2int_rn(stz);
    }
}

// probe_brick, raytracing_on_cpu.rs:256-312. kind: 0 empty, 1 parted, 2 solid
__device__ __forceinline__ bool probe_brick(const DeviceTree& t, const RayConst& r, float& px, float& py, float& pz,
                                            uint32_t kind, uint32_t slot, float bx, float by, float bz, float bsize,
                                            float inv_size, TraceResult& out) {
    if (kind == BK_EMPTY) return false;
    if (kind == BK_SOLID) {
        out.palette_value = slot;
        out.px = px; out.py = py; out.pz = pz;
        out.bx = bx; out.by = by; out.bz = bz; out.bsize = bsize;
        return true;
    }
    int hx, hy, hz;
    const int flat = traverse_brick(t, r, px, py, pz, slot, bx, by, bz, bsize, inv_size, hx, hy, hz);
    if (flat < 0) return false;
    out.palette_value = __ldg(t.voxels + ((size_t)slot << (3 * t.brick_shift)) + flat);
    out.px = px; out.py = py; out.pz = pz;
    // hit_bounds: min + idx * size / dim (the division is by a power of two), size / dim
    const float inv_dim = t.inv_brick_dim;
    out.bx = bx + ((float)hx * bsize) * inv_dim;
    out.by = by + ((float)hy * bsize) * inv_dim;
    out.bz = bz + ((float)hz * bsize) * inv_dim;
    out.bsize = bsize * inv_dim;
    return true;
}

// Octree::get_by_ray -> get_by_ray_at_lod(ray, f32::MAX), raytracing_on_cpu.rs:316-565 (MIP maps off: :369-386 dead)
__device__ __forceinline__ bool trace_ray(const DeviceTree& t, const RayConst& r, TraceResult& out) {
    const float tree_size = (float)t.tree_size;
    float px, py, pz;
    uint32_t target_octant;
    {
        // Cube::intersect_ray on the root cube, spatial/raytracing/mod.rs:32-61
        const float t1 = (0.0f - r.ox) / r.dx, t2 = (tree_size - r.ox) / r.dx;
        const float t3 = (0.0f - r.oy) / r.dy, t4 = (tree_size - r.oy) / r.dy;
        const float t5 = (0.0f - r.oz) / r.dz, t6 = (tree_size - r.oz) / r.dz;
        const float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
        const float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
        if (tmax < 0.0f || tmin > tmax) {
            out.palette_value = NIL;
            return false;
        }
        const float d = (tmin < 0.0f) ? 0.0f : tmin;
        px = r.ox + r.dx * d;
        py = r.oy + r.dy * d;
        pz = r.oz + r.dz * d;
        target_octant = hash_region(px, py, pz, tree_size * 0.5f);
    }
    // NodeStack<u32, 4>, raytracing_on_cpu.rs:20-82: a ring buffer that overwrites its oldest entry
    // Held as a 4-deep shift register (s0 = newest): pushing drops the oldest entry, popping removes the newest,
    // which is exactly what the ring buffer does; entries beyond `count` are never read.
    uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u;
    uint32_t count = 0;
    uint32_t cur = 0;
    float bx = 0.0f, by = 0.0f, bz = 0.0f, bsize = tree_size;
    float binv = t.inv_tree_size;  // 1 / bsize, exact (powers of two), tracked alongside bsize

    while (target_octant != OOB_OCTANT) {
        cur = 0;
        bx = by = bz = 0.0f;
        bsize = tree_size;
        binv = t.inv_tree_size;
        s3 = s2; s2 = s1; s1 = s0; s0 = 0u;
        count = min(count + 1u, 4u);
        while (count != 0u) {
            // cur == stack[head] at this point: one 16-byte load serves occupancy bits and node kind
            const uint4 hd = __ldg(reinterpret_cast<const uint4*>(t.node_head) + cur);
            const uint32_t oc_lo = hd.x, oc_hi = hd.y, meta = hd.z;
            const uint32_t kind = meta & 3u;
            bool backtrack = (kind == NK_UNIFORM);
            if (target_octant != OOB_OCTANT) {
                if (kind == NK_UNIFORM) {
                    if (probe_brick(t, r, px, py, pz, (meta >> 2) & 3u, hd.w, bx, by, bz, bsize, binv, out)) return true;
                } else if (kind == NK_LEAF) {
                    const uint32_t bkind = (meta >> (2u + 2u * target_octant)) & 3u;
                    if (bkind != BK_EMPTY) {
                        const float hs = bsize * 0.5f;
                        const uint32_t slot = __ldg(t.node_slot + (size_t)cur * 8u + target_octant);
                        if (probe_brick(t, r, px, py, pz, bkind, slot, bx + (float)(target_octant & 1u) * hs,
                                        by + (float)((target_octant >> 2) & 1u) * hs,
                                        bz + (float)((target_octant >> 1) & 1u) * hs, hs, binv * 2.0f, out))
                            return true;
                    }
                }
            }
            // position inside the node in 4x4x4 bitmap cells (:425-436)
            float bpx = rust_clamp(((px - bx) * 4.0f) * binv, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
            float bpy = rust_clamp(((py - by) * 4.0f) * binv, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
            float bpz = rust_clamp(((pz - bz) * 4.0f) * binv, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
            if (backtrack || target_octant == OOB_OCTANT || (oc_lo | oc_hi) == 0u ||
                !ray_may_hit(oc_lo, oc_hi, bitmap_coord(bpx), bitmap_coord(bpy), bitmap_coord(bpz), r.dirbits)) {
                // POP (:445-474)
                count -= 1u;
                s0 = s1; s1 = s2; s2 = s3;
                if (count != 0u) {
                    cur = s0;
                    const float twice = bsize * 2.0f;
                    const float inv_twice = binv * 0.5f;
                    // parent min = min - min % (2*size) (:452-456). min is a non-negative exact multiple of size, so
                    // this equals floor(min / 2size) * 2size, all steps exact in f32.
                    const float pbx = floorf(bx * inv_twice) * twice, pby = floorf(by * inv_twice) * twice,
                                pbz = floorf(bz * inv_twice) * twice;
                    const float half = bsize * 0.5f;
                    const uint32_t from = hash_region((bx + half) - pbx, (by + half) - pby, (bz + half) - pbz, bsize);
                    float stx, sty, stz;
                    dda_step(r, px, py, pz, bx, by, bz, bsize, stx, sty, stz);
                    target_octant = step_octant(from, stx, sty, stz);
                    bsize = twice;
                    binv = inv_twice;
                    bx = pbx; by = pby; bz = pbz;
                }
                continue;
            }
            const float hs = bsize * 0.5f;
            float tbx = bx + (float)(target_octant & 1u) * hs;
            float tby = by + (float)((target_octant >> 2) & 1u) * hs;
            float tbz = bz + (float)((target_octant >> 1) & 1u) * hs;
            // NodeChildren::child(): only Internal nodes carry child keys (node.rs:49-54)
            uint32_t child = (kind == NK_INTERNAL) ? __ldg(t.node_slot + (size_t)cur * 8u + target_octant) : NIL;
            if (child != NIL && octant_occupied(oc_lo, oc_hi, target_octant)) {
                // PUSH (:484-492)
                cur = child;
                bx = tbx; by = tby; bz = tbz;
                bsize = hs;
                binv = binv * 2.0f;
                target_octant = hash_region(px - bx, py - by, pz - bz, hs * 0.5f);
                s3 = s2; s2 = s1; s1 = s0; s0 = child;
                count = min(count + 1u, 4u);
            } else {
                // ADVANCE (:497-544)
                for (;;) {
                    float stx, sty, stz;
                    dda_step(r, px, py, pz, tbx, tby, tbz, hs, stx, sty, stz);
                    target_octant = step_octant(target_octant, stx, sty, stz);
                    if (target_octant == OOB_OCTANT) break;
                    tbx = bx + (float)(target_octant & 1u) * hs;
                    tby = by + (float)((target_octant >> 2) & 1u) * hs;
                    tbz = bz + (float)((target_octant >> 1) & 1u) * hs;
                    // (sic) 4/size cells per sibling step, not 2
                    bpx = bpx + (stx * 4.0f) * binv;
                    bpy = bpy + (sty * 4.0f) * binv;
                    bpz = bpz + (stz * 4.0f) * binv;
                    if (kind == NK_INTERNAL) {
                        child = __ldg(t.node_slot + (size_t)cur * 8u + target_octant);
                        if (child != NIL && octant_occupied(oc_lo, oc_hi, target_octant) &&
                            ray_may_hit(oc_lo, oc_hi, bitmap_coord(bpx), bitmap_coord(bpy), bitmap_coord(bpz), r.dirbits))
                            break;
                    } else if (kind == NK_LEAF) {
                        if (((meta >> (2u + 2u * target_octant)) & 3u) != BK_EMPTY) break;
                    }
                }
            }
        }
        // restart from the root after a 0.1 nudge (:548-562)
        px = px + r.dx * 0.1f;
        py = py + r.dy * 0.1f;
        pz = pz + r.dz * 0.1f;
        if (px < tree_size && py < tree_size && pz < tree_size && px > 0.0f && py > 0.0f && pz > 0.0f)
            target_octant = hash_region(px, py, pz, tree_size * 0.5f);
        else
            target_octant = OOB_OCTANT;
    }
    out.palette_value = NIL;
    return false;
}

// cube_impact_normal, spatial/raytracing/mod.rs:106-134
__device__ __forceinline__ void impact_normal(const TraceResult& h, float& nx, float& ny, float& nz) {
    const float half = h.bsize * 0.5f;
    const float mx = (h.bx + half) - h.px, my = (h.by + half) - h.py, mz = (h.bz + half) - h.pz;
    const float mc = fmaxf(fmaxf(fabsf(mx), fabsf(my)), fabsf(mz));
    const float ax = (fabsf(mx) == mc) ? -mx : 0.0f;
    const float ay = (fabsf(my) == mc) ? -my : 0.0f;
    const float az = (fabsf(mz) == mc) ? -mz : 0.0f;
    const float len = sqrtf((ax * ax) + (ay * ay) + (az * az));
    nx = ax / len;
    ny = ay / len;
    nz = az / len;
}

}  // namespace svx
