// Device-side primary-ray traversal of the brick-leaf sparse voxel octree: Octree::get_by_ray of the reference
// (src/raytracing/raytracing_on_cpu.rs:316-565) restated for sm_100a. NOT a port of the WGSL shader (which is a
// different, lossy variant - SURVEY §2.2); results must equal the CPU get_by_ray bit for bit.
//
// Numerics: compiled with -fmad=false (Rust never contracts to FMA), IEEE division / sqrt (nvcc defaults
// -prec-div=true -prec-sqrt=true, -ftz=false). Divisions by node / brick sizes are by exact powers of two and are
// written as multiplications by the exact reciprocal: x / 2^k and x * 2^-k are the same real number, both correctly
// rounded, hence the same bits. Node bounds are exact small integers in f32.
//
// The reference's small look-up tables (src/spatial/lut.rs) are replaced by closed forms of their generator logic
// (lut.rs:12-152); the 4 KB RAY_TO_NODE_OCCUPANCY_BITMASK_LUT is read as a table (DeviceTree::ray_lut, regenerated on
// the host, L1-resident): measured, one 8-byte load beats its ~33-instruction closed form. The start-up self-test
// (capi.cu: run_selftest) checks every table entry against the closed forms.
// `file:line` citations are relative to the reference checkout.
#pragma once
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <cuda_runtime.h>
// SVX_HOST_MIRROR: this header compiled by the HOST compiler in tests/host_mirror (test infrastructure, never part of the
// library): the PTX fragments below get plain C++ equivalents so that the kernels' logic - the transformed arithmetic,
// the mirrored brick walk, the parent-index stack, the crawl fast-forward - can be checked against the oracle without a
// GPU. Device compilation is unaffected (SVX_HOST_MIRROR is never defined there).

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

// The same test through the table itself (DeviceTree::ray_lut, [direction octant][cell] {lo, hi}): what the reference
// does (`RAY_TO_NODE_OCCUPANCY_BITMASK_LUT[flat_pos][direction_lut_index] & occupied_bits`, raytracing_on_cpu.rs:443),
// one L1-resident 8-byte load instead of ~30 integer instructions. SVX_RAY_LUT=0 builds the closed form into the kernels.
#ifndef SVX_RAY_LUT
#define SVX_RAY_LUT 1
#endif
__device__ __forceinline__ bool ray_may_hit_node(const DeviceTree& t, uint32_t oc_lo, uint32_t oc_hi, uint32_t cx, uint32_t cy,
                                                 uint32_t cz, uint32_t dirbits) {
#if SVX_RAY_LUT
    const uint2 m = __ldg(t.ray_lut + (((dirbits & 7u) << 6) + cx + (cy << 2) + (cz << 4)));
    return ((oc_lo & m.x) | (oc_hi & m.y)) != 0u;
#else
    return ray_may_hit(oc_lo, oc_hi, cx, cy, cz, dirbits & 7u);
#endif
}

// step_octant, spatial/raytracing/mod.rs:68-80 with OCTANT_STEP_RESULT_LUT (generate_octant_step_result_lut,
// lut.rs:91-137): move one octant along each stepped axis, OOB when leaving the 2x2x2 block. The step arrives as three
// "this axis stepped" predicates; its per-axis sign is the ray's (`signum` of the direction, dda_step_to_next_sibling),
// given as `posbits` in the octant's own bit layout (bit0 x, bit1 z, bit2 y; set = the ray moves up that axis). A stepped
// axis leaves the block exactly when the octant already sits on the side the ray moves towards, i.e. when its octant bit
// equals its posbit; otherwise the bit flips.
__device__ __forceinline__ uint32_t step_octant(uint32_t octant, bool sx, bool sy, bool sz, uint32_t posbits) {
    const uint32_t stepped = (sx ? 1u : 0u) | (sz ? 2u : 0u) | (sy ? 4u : 0u);
    if ((stepped & ~(octant ^ posbits)) != 0u) return OOB_OCTANT;
    return octant ^ stepped;
}

// The 64-byte record of node i (gpu_tree.hpp): [0] head, [1..2] the eight slots, [3] the bounds
__device__ __forceinline__ const uint4* node_record(const DeviceTree& t, uint32_t i) { return t.node_rec + (size_t)i * 4u; }
__device__ __forceinline__ uint4 node_head_of(const uint4* rec) { return __ldg(rec); }
__device__ __forceinline__ uint32_t node_slot_of(const uint4* rec, uint32_t octant) {
    return __ldg(reinterpret_cast<const uint32_t*>(rec + 1) + octant);
}
__device__ __forceinline__ float4 node_bounds_of(const uint4* rec) { return __ldg(reinterpret_cast<const float4*>(rec + 3)); }

struct RayConst {
    float ox, oy, oz;      // origin
    float dx, dy, dz;      // direction
    float sfx, sfy, sfz;   // get_dda_scale_factors, raytracing_on_cpu.rs:99-112
    bool negx, negy, negz; // sign bit of the direction: f32::signum is -1.0 (also for -0.0), else +1.0
    int isx, isy, isz;     // the same as integers
    // bits 0-2: hash_direction (spatial/math/mod.rs:22-26), the ray-to-node table index. bits 3-5: the step signs in the
    // octant bit layout, !negx | !negz << 1 | !negy << 2 (step_octant). They differ for components in (-2^-25, -0]
    // (`1 + d >= 1` rounds those to "positive"), so both are kept - in one register, or the compiler rebuilds the second
    // from the sign predicates at every use.
    uint32_t dirbits;
};

// Blackwell's packed f32 arithmetic (sm_100: add / sub / mul.rn.f32x2 -> FADD2 / FMUL2, two IEEE round-to-nearest results per
// issued instruction) for the x and y components of the DDA's vector operations; z stays scalar. Each half is the same
// correctly rounded operation as the scalar instruction, so results are unchanged. A packed multiply must never feed a
// packed add directly: ptxas 12.9 contracts that pair into FFMA2 even for .rn operands and under --fmad=false (one
// rounding instead of two); sums of products therefore use scalar adds, and the build refuses a library with FFMA2 in it
// (build.py: check_no_packed_fma).
#ifndef SVX_PACKED_DDA
#define SVX_PACKED_DDA 1
#endif
// 1: the voxel loop loads its 32-voxel occupancy word on every step (one L1-resident load, 34 issued instructions per step).
// 0: it keeps the word in a register and reloads when the word index changes (36 without / 41 with a reload). Measured on
// B200: the unconditional load is 3-9 % faster per frame - the kernel is bound by instruction issue, not by L1.
// 1: one instantiation of the brick walk serves UniformLeaf nodes and the octants of Leaf nodes (smaller kernel, the two
// kinds of lanes converge in the walk); 0: one instantiation per kind
#ifndef SVX_SINGLE_PROBE_SITE
#define SVX_SINGLE_PROBE_SITE 1
#endif
#ifndef SVX_BRICK_WORD_ALWAYS
#define SVX_BRICK_WORD_ALWAYS 1
#endif
#ifndef SVX_HOST_MIRROR
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); return v; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t v; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(a), "l"(b)); return v; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t v; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(a), "l"(b)); return v; }
#else  // tests/host_mirror: the same helpers for the host compiler (each half is one IEEE operation, as on the device)
inline uint64_t pack2(float lo, float hi) { uint32_t a, b; std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4); return (uint64_t)a | ((uint64_t)b << 32); }
inline void unpack2(uint64_t v, float& lo, float& hi) { const uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32); std::memcpy(&lo, &a, 4); std::memcpy(&hi, &b, 4); }
inline uint64_t sub2(uint64_t a, uint64_t b) { float al, ah, bl, bh; unpack2(a, al, ah); unpack2(b, bl, bh); return pack2(al - bl, ah - bh); }
inline uint64_t mul2(uint64_t a, uint64_t b) { float al, ah, bl, bh; unpack2(a, al, ah); unpack2(b, bl, bh); return pack2(al * bl, ah * bh); }
#endif

// dda_step_to_next_sibling, raytracing_on_cpu.rs:124-152.
//   steps_needed = size * signum.max(0.) - signum * (p - min)      (signum = +-1.0)
// is `size - (p - min)` for signum +1 (x*1 and 1*x are exact) and `0 - (-(p - min))` = p - min for signum -1; the
// only bit that can differ is the sign of a zero, which the following `.abs()` removes. A NaN direction component
// makes its scale factor NaN, hence d NaN, in both forms. Outputs: which axes stepped (min_step == d_axis).
__device__ __forceinline__ void dda_step(const RayConst& r, float& px, float& py, float& pz, float bx, float by,
                                         float bz, float bsize, bool& sx, bool& sy, bool& sz) {
#if SVX_PACKED_DDA
    float dfx, dfy;
    unpack2(sub2(pack2(px, py), pack2(bx, by)), dfx, dfy);
    const float dfz = pz - bz;
    const float nx = r.negx ? dfx : (bsize - dfx);
    const float ny = r.negy ? dfy : (bsize - dfy);
    const float nz = r.negz ? dfz : (bsize - dfz);
    float tx, ty;
    unpack2(mul2(pack2(nx, ny), pack2(r.sfx, r.sfy)), tx, ty);
    const float d_x = fabsf(tx), d_y = fabsf(ty), d_z = fabsf(nz * r.sfz);
    const float m = fminf(fminf(d_x, d_y), d_z);
    float mx, my;
    unpack2(mul2(pack2(r.dx, r.dy), pack2(m, m)), mx, my);  // the sums stay scalar: see the note on FFMA2 above
    px = px + mx;
    py = py + my;
    pz = pz + r.dz * m;
#else
    const float dfx = px - bx, dfy = py - by, dfz = pz - bz;
    const float nx = r.negx ? dfx : (bsize - dfx);
    const float ny = r.negy ? dfy : (bsize - dfy);
    const float nz = r.negz ? dfz : (bsize - dfz);
    const float d_x = fabsf(nx * r.sfx);
    const float d_y = fabsf(ny * r.sfy);
    const float d_z = fabsf(nz * r.sfz);
    const float m = fminf(fminf(d_x, d_y), d_z);
    px = px + r.dx * m;
    py = py + r.dy * m;
    pz = pz + r.dz * m;
#endif
    sx = (m == d_x);
    sy = (m == d_y);
    sz = (m == d_z);
}

// The same step against a cube whose size stays fixed over several steps (the sibling walk of ADVANCE): steps_needed in the
// sign-free form (p - min) - off with off = 0 along a descending axis and `size` along an ascending one, precomputed by
// the caller. RN(d - size) = -RN(size - d) and d - 0 = d, and only the magnitude is used (`.abs()`, :139-143), so this is
// dda_step's value for either sign without the per-axis select.
struct DdaStep {
    float m, d_x, d_y, d_z;  // min_step and the per-axis distances it was chosen from: axis a stepped iff m == d_a
};
__device__ __forceinline__ DdaStep dda_step_off(const RayConst& r, float& px, float& py, float& pz, uint64_t bxy, float bz,
                                                uint64_t offxy, float offz, bool& sx, bool& sy, bool& sz) {
    float tx, ty;
    unpack2(mul2(sub2(sub2(pack2(px, py), bxy), offxy), pack2(r.sfx, r.sfy)), tx, ty);
    const float d_x = fabsf(tx), d_y = fabsf(ty), d_z = fabsf(((pz - bz) - offz) * r.sfz);
    const float m = fminf(fminf(d_x, d_y), d_z);
    float mx, my;
    unpack2(mul2(pack2(r.dx, r.dy), pack2(m, m)), mx, my);
    px = px + mx;
    py = py + my;
    pz = pz + r.dz * m;
    sx = (m == d_x);
    sy = (m == d_y);
    sz = (m == d_z);
    return DdaStep{m, d_x, d_y, d_z};
}

// `if (m == d) { a += da; b += db; }` as one compare and two predicated additions
__device__ __forceinline__ void add_both_if_equal(float m, float d, float& a, float da, float& b, float db) {
#ifndef SVX_HOST_MIRROR
    asm("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %2, %3;\n\t@p add.rn.f32 %0, %0, %4;\n\t@p add.rn.f32 %1, %1, %5;\n\t}"
        : "+f"(a), "+f"(b) : "f"(m), "f"(d), "f"(da), "f"(db));
#else
    if (m == d) { a = a + da; b = b + db; }
#endif
}

// `negative ? -v : v` for v >= +0 as one logic instruction: v with the sign bit of `direction` (f32::signum's sign, also for
// -0.0). And `negative ? 0 : v` from that: max(+-v, 0).
__device__ __forceinline__ float with_sign_of(float v, float direction) {
    return __uint_as_float(__float_as_uint(v) | (__float_as_uint(direction) & 0x80000000u));
}

// The part of ray_setup that does not divide
__device__ __forceinline__ void ray_setup_signs(RayConst& r);

// Everything get_by_ray derives from the direction before the loop (raytracing_on_cpu.rs:331-332)
__device__ __forceinline__ void ray_setup(RayConst& r) {
    auto sq = [](float v) { return v * v; };  // `.powf(2.)` == x*x
    r.sfx = sqrtf(1.0f + sq(r.dz / r.dx) + sq(r.dy / r.dx));
    r.sfy = sqrtf(sq(r.dx / r.dy) + 1.0f + sq(r.dz / r.dy));
    r.sfz = sqrtf((sq(r.dx / r.dz) + 1.0f) + sq(r.dy / r.dz));
    ray_setup_signs(r);
}
__device__ __forceinline__ void ray_setup_signs(RayConst& r) {
    r.negx = signbit(r.dx);
    r.negy = signbit(r.dy);
    r.negz = signbit(r.dz);
    r.isx = r.negx ? -1 : 1;
    r.isy = r.negy ? -1 : 1;
    r.isz = r.negz ? -1 : 1;
    r.dirbits = hash_region(1.0f + r.dx, 1.0f + r.dy, 1.0f + r.dz, 1.0f) | (r.negx ? 0u : 8u) | (r.negz ? 0u : 16u) | (r.negy ? 0u : 32u);
    // opaque from here on: the value lives in its register instead of being rebuilt (three additions, six compares,
    // selects) inside the node and sibling loops, which is what the compiler otherwise prefers
#ifndef SVX_HOST_MIRROR
    asm volatile("" : "+r"(r.dirbits));
#endif
}

// `(v as i32).clamp(0, dim-1)`: cvt.rzi saturates and maps NaN to 0 like Rust's `as`
__device__ __forceinline__ int clamp_index(float v, int dim) { return min(max(__float2int_rz(v), 0), dim - 1); }
// `v.floor() as usize` for the 4x4x4 bitmap position; the reference bounds-panics above 3, we clamp
__device__ __forceinline__ uint32_t bitmap_coord(float v) { return (uint32_t)min(max(__float2int_rd(v), 0), 3); }
// The same for a value that just went through rust_clamp(v, 1e-5, 4 - 1e-5) (:430-434): it is in [1e-5, 3.99999] or NaN,
// so the floor is already 0..3 (cvt maps NaN to 0) and the integer clamp is dead code
__device__ __forceinline__ uint32_t bitmap_coord_of_clamped(float v) { return (uint32_t)__float2int_rd(v); }

// Brick geometry. BS >= 0: the brick dimension 2^BS is a compile-time constant of the kernel instantiation (the host
// picks the instantiation that matches DeviceTree::brick_shift, kernels.cu: launch_render), so strides, masks and the
// reciprocal are immediates instead of constant-bank loads and registers inside the voxel loop. BS < 0: read from the tree.
template <int BS> __device__ __forceinline__ uint32_t brick_shift_of(const DeviceTree& t) { if constexpr (BS >= 0) return (uint32_t)BS; else return t.brick_shift; }
template <int BS> __device__ __forceinline__ uint32_t brick_dim_of(const DeviceTree& t) { if constexpr (BS >= 0) return 1u << BS; else return t.brick_dim; }
template <int BS> __device__ __forceinline__ uint32_t brick_dim_sq_of(const DeviceTree& t) { if constexpr (BS >= 0) return 1u << (2 * BS); else return t.brick_dim_sq; }
template <int BS> __device__ __forceinline__ uint32_t bit_words_of(const DeviceTree& t) { if constexpr (BS >= 0) return ((1u << (3 * BS)) + 31u) / 32u; else return t.bit_words; }
template <int BS> __device__ __forceinline__ float inv_brick_dim_of(const DeviceTree& t) { if constexpr (BS >= 0) return 1.0f / (float)(1u << BS); else return t.inv_brick_dim; }

// &base[i] formed by one multiply-add on the address (mad.wide.u32); written in PTX because the compiler otherwise
// distributes the scaling over the index expression and spends four instructions on it
__device__ __forceinline__ const uint32_t* word_address(const uint32_t* base, uint32_t i) {
#ifndef SVX_HOST_MIRROR
    uint64_t a;
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(a) : "r"(i), "l"(reinterpret_cast<uint64_t>(base)));
    return reinterpret_cast<const uint32_t*>(a);
#else
    return base + i;
#endif
}

// traverse_brick, raytracing_on_cpu.rs:156-252. Walks the occupancy bit-brick (1 bit per voxel, set = not empty) and
// returns the flat index (flat_projection, math/mod.rs:35-37) of the first non-empty voxel or -1.
// The loop carries only what a step needs: the flat index (it addresses the bit and, on a hit, gives the voxel index
// back) and the min corner of the current cell; the 32-voxel occupancy word is loaded on every step (L1-resident, see
// SVX_BRICK_WORD_ALWAYS). The reference's bounds check
// on the integer index (:193-203) is done on that corner instead: corners are exact multiples of `unit` (integers),
// every step moves a stepped axis by exactly one cell, so the walk has left the brick exactly when a corner equals
// the first corner outside (min - unit going down, min + size going up).
template <int BS>
__device__ __forceinline__ int traverse_brick(const DeviceTree& t, const RayConst& r, float& px, float& py, float& pz,
                                              uint32_t brick, float bx, float by, float bz, float bsize, float inv_size) {
    const int dim = (int)brick_dim_of<BS>(t);
    const float fdim = (float)dim;
    // `(p - min) * dim / size` (:167-170): dim and size are powers of two, so the two scalings are one by dim / size (exact)
    const float to_cells = fdim * inv_size;
    const int ix = clamp_index((px - bx) * to_cells, dim);
    const int iy = clamp_index((py - by) * to_cells, dim);
    const int iz = clamp_index((pz - bz) * to_cells, dim);
    const float unit = bsize * inv_brick_dim_of<BS>(t);  // size / dim, exact: both powers of two
    float cx = bx + (float)ix * unit, cy = by + (float)iy * unit, cz = bz + (float)iz * unit;
    // `current_bounds.min_position += step * brick_unit`: step is +-1.0 or 0.0, so the addend is +-unit or +0
    const float ux = with_sign_of(unit, r.dx), uy = with_sign_of(unit, r.dy), uz = with_sign_of(unit, r.dz);
    const float ex = r.negx ? bx - unit : bx + bsize, ey = r.negy ? by - unit : by + bsize, ez = r.negz ? bz - unit : bz + bsize;
    const uint32_t sh = brick_shift_of<BS>(t);
    // flat_projection(ix, iy, iz) kept incrementally, like the reference's current_flat_index (:205-207) - but in
    // MIRRORED coordinates: along an axis the ray descends, the loop counts j = dim-1 - i = i ^ (dim-1) instead of i, so
    // every step adds +1 / +dim / +dim^2 (uniform values, no per-ray signed strides to keep or rebuild in the loop) and
    // the real flat index is `mirrored ^ flip` with one per-brick mask. A walk that leaves the brick is caught by the
    // corner test below before the (then meaningless) index is used again.
    const uint32_t dmask = (uint32_t)dim - 1u;
    const uint32_t flip = (r.negx ? dmask : 0u) | (r.negy ? dmask << sh : 0u) | (r.negz ? dmask << (2 * sh) : 0u);
    uint32_t mirrored = ((uint32_t)ix + ((uint32_t)iy << sh) + ((uint32_t)iz << (2 * sh))) ^ flip;
    const uint32_t base = brick * bit_words_of<BS>(t);  // word offset of this brick's bits; all bit words fit 32 bits (gpu_tree.cpp)
    uint32_t word = 0u;
#if SVX_PACKED_DDA
    // dda_step with the x and y components packed. `steps_needed` in the sign-free form (p - corner) - off, off = 0 along a
    // descending axis and `unit` along an ascending one: RN(d - unit) = -RN(unit - d) and d - 0 = d, and only the magnitude
    // is used (the `.abs()` of :139-143), so this is dda_step's value for either sign without a per-axis select.
    const uint64_t offxy = pack2(fmaxf(ux, 0.0f), fmaxf(uy, 0.0f)), sfxy = pack2(r.sfx, r.sfy), dxy = pack2(r.dx, r.dy);  // negative ? 0 : unit
    float offz = fmaxf(uz, 0.0f);
#ifndef SVX_HOST_MIRROR
    asm volatile("" : "+f"(offz));  // a loop constant in a register (otherwise rebuilt from `unit` and the sign on every step)
#endif
    uint64_t pxy = pack2(px, py);
    // The loop counts the COMPLEMENT of the flat index (mirrored ^ ~flip): the voxel's bit is then moved to the sign position
    // by a left shift of (~flat & 31) = 31 - (flat & 31), a one-instruction test.
    uint32_t nflip = ~flip;
#ifndef SVX_HOST_MIRROR
    asm volatile("" : "+r"(nflip));  // keeps the complement a loop constant (the compiler would rather complement on every step)
#endif
    // word (flat >> 5) of the brick = bits[(nflat >> 5) ^ 0x07FFFFFF]
    const uint32_t* bits = t.brick_bits + base;
#ifndef SVX_HOST_MIRROR
    asm volatile("" : "+l"(bits));  // the address of the brick's words stays in a register pair
#endif
#if !SVX_BRICK_WORD_ALWAYS
    uint32_t nword_index = 0u;  // no complemented word index of a brick is 0 (their upper bits are set)
#endif
    uint32_t nflat;
    for (;;) {
        nflat = mirrored ^ nflip;
#if SVX_BRICK_WORD_ALWAYS
        // one L1-resident load per step instead of "same word as before?" bookkeeping: fewer issued instructions
        word = __ldg(word_address(bits, (nflat >> 5) ^ 0x07FFFFFFu));
#else
        if ((nflat >> 5) != nword_index) {
            nword_index = nflat >> 5;
            word = __ldg(word_address(bits, nword_index ^ 0x07FFFFFFu));
        }
#endif
        if ((int)(word << (nflat & 31u)) < 0) break;
        float tx, ty;
        unpack2(mul2(sub2(sub2(pxy, pack2(cx, cy)), offxy), sfxy), tx, ty);
        const float tz = ((pz - cz) - offz) * r.sfz;
        const float d_x = fabsf(tx), d_y = fabsf(ty), d_z = fabsf(tz);
        const float m = fminf(fminf(d_x, d_y), d_z);
        float mx, my, qx, qy;
        unpack2(mul2(dxy, pack2(m, m)), mx, my);
        unpack2(pxy, qx, qy);
        pxy = pack2(qx + mx, qy + my);
        pz = pz + r.dz * m;
        // `if (m == d) { mirrored += stride; corner += u; }` per axis, as predicated instructions
#ifndef SVX_HOST_MIRROR
        asm("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %2, %3;\n\t@p add.u32 %0, %0, %4;\n\t@p add.rn.f32 %1, %1, %5;\n\t}"
            : "+r"(mirrored), "+f"(cx) : "f"(m), "f"(d_x), "r"(1u), "f"(ux));
        asm("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %2, %3;\n\t@p add.u32 %0, %0, %4;\n\t@p add.rn.f32 %1, %1, %5;\n\t}"
            : "+r"(mirrored), "+f"(cy) : "f"(m), "f"(d_y), "r"(brick_dim_of<BS>(t)), "f"(uy));
        asm("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %2, %3;\n\t@p add.u32 %0, %0, %4;\n\t@p add.rn.f32 %1, %1, %5;\n\t}"
            : "+r"(mirrored), "+f"(cz) : "f"(m), "f"(d_z), "r"(brick_dim_sq_of<BS>(t)), "f"(uz));
#else
        if (m == d_x) { mirrored += 1u; cx = cx + ux; }
        if (m == d_y) { mirrored += brick_dim_of<BS>(t); cy = cy + uy; }
        if (m == d_z) { mirrored += brick_dim_sq_of<BS>(t); cz = cz + uz; }
#endif
        if (cx == ex || cy == ey || cz == ez) break;
    }
    unpack2(pxy, px, py);
    // the walk ended on a set bit (corners strictly inside the brick) or by leaving it (a corner on the first plane outside)
    return (cx == ex || cy == ey || cz == ez) ? -1 : (int)~nflat;
#else
    uint32_t word_index = 0xFFFFFFFFu;
    for (;;) {
        const uint32_t flat = mirrored ^ flip;
        if ((flat >> 5) != word_index) {
            word_index = flat >> 5;
            word = __ldg(t.brick_bits + (base + word_index));
        }
        if ((word >> (flat & 31u)) & 1u) return (int)flat;
        bool sx, sy, sz;
        dda_step(r, px, py, pz, cx, cy, cz, unit, sx, sy, sz);
        if (sx) { mirrored += 1u; cx = cx + ux; }
        if (sy) { mirrored += brick_dim_of<BS>(t); cy = cy + uy; }
        if (sz) { mirrored += brick_dim_sq_of<BS>(t); cz = cz + uz; }
        if (cx == ex || cy == ey || cz == ez) return -1;
    }
#endif
}

// probe_brick, raytracing_on_cpu.rs:256-312. kind: 0 empty, 1 parted, 2 solid
template <int BS>
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
    const int flat = traverse_brick<BS>(t, r, px, py, pz, slot, bx, by, bz, bsize, inv_size);
    if (flat < 0) return false;
    const uint32_t dmask = brick_dim_of<BS>(t) - 1u, sh = brick_shift_of<BS>(t);
    const int hx = (int)((uint32_t)flat & dmask), hy = (int)(((uint32_t)flat >> sh) & dmask), hz = (int)((uint32_t)flat >> (2u * sh));
    out.palette_value = __ldg(t.voxels + ((size_t)slot << (3u * sh)) + flat);
    out.px = px; out.py = py; out.pz = pz;
    // hit_bounds: min + idx * size / dim (the division is by a power of two), size / dim
    const float inv_dim = inv_brick_dim_of<BS>(t);
    out.bx = bx + ((float)hx * bsize) * inv_dim;
    out.by = by + ((float)hy * bsize) * inv_dim;
    out.bz = bz + ((float)hz * bsize) * inv_dim;
    out.bsize = bsize * inv_dim;
    return true;
}

// Conservative "this ray certainly misses the root cube" test on APPROXIMATE arithmetic (MUFU reciprocals, no IEEE
// division). `dx,dy,dz` may be any direction within a few ulp of the exactly normalised one. With a_i the exact
// numerators (bound - origin, the same single subtraction the exact path performs), every approximate t_i is within
// 2^-19 relative of the t_i the exact slab test (root_entry) computes; min/max are 1-Lipschitz, so tmin/tmax are
// within 2^-19 * M, M = max|t_i|. The test only answers "miss" with a 2^-14 * M margin (32x slack) and answers
// "don't know" for anything non-finite or nearly axis-parallel; those rays take the exact path. It can therefore
// never change a result, only skip work.
__device__ __forceinline__ bool certain_root_miss(float ox, float oy, float oz, float dx, float dy, float dz, float size) {
    const float dmin = fminf(fminf(fabsf(dx), fabsf(dy)), fabsf(dz));
    if (!(dmin > 1e-20f)) return false;
    const float rx = __fdividef(1.0f, dx), ry = __fdividef(1.0f, dy), rz = __fdividef(1.0f, dz);
    const float t1 = (0.0f - ox) * rx, t2 = (size - ox) * rx;
    const float t3 = (0.0f - oy) * ry, t4 = (size - oy) * ry;
    const float t5 = (0.0f - oz) * rz, t6 = (size - oz) * rz;
    const float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    const float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    const float m = fmaxf(fmaxf(fmaxf(fabsf(t1), fabsf(t2)), fmaxf(fabsf(t3), fabsf(t4))), fmaxf(fabsf(t5), fabsf(t6)));
    if (!(m < 1e30f)) return false;
    const float e = m * 6.103515625e-5f;  // 2^-14
    return (tmax < -e) || ((tmin - tmax) > 2.0f * e);
}

// Not part of the reference: an entry point that is not finite (origin or direction at the ends of the f32 range, NaN
// rays) makes every DDA distance NaN, no axis ever steps and the walks below would spin forever - the reference trips its
// debug assertions or hangs on its CPU thread there. The kernels report a miss instead. x - x is 0 for finite x, NaN otherwise.
__device__ __forceinline__ bool all_finite(float x, float y, float z) { return ((x - x) + (y - y)) + (z - z) == 0.0f; }

// Cube::intersect_ray on the root cube + the entry point and octant (spatial/raytracing/mod.rs:32-61,
// raytracing_on_cpu.rs:335-348). Needs only origin and direction of `r`. Returns false when the ray misses.
__device__ __forceinline__ bool root_entry(const RayConst& r, float tree_size, float& px, float& py, float& pz,
                                           uint32_t& target_octant) {
    const float t1 = (0.0f - r.ox) / r.dx, t2 = (tree_size - r.ox) / r.dx;
    const float t3 = (0.0f - r.oy) / r.dy, t4 = (tree_size - r.oy) / r.dy;
    const float t5 = (0.0f - r.oz) / r.dz, t6 = (tree_size - r.oz) / r.dz;
    const float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    const float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    if (tmax < 0.0f || tmin > tmax) return false;
    const float d = (tmin < 0.0f) ? 0.0f : tmin;  // impact_distance.unwrap_or(0.)
    px = r.ox + r.dx * d;
    py = r.oy + r.dy * d;
    pz = r.oz + r.dz * d;
    if (!all_finite(px, py, pz)) return false;
    target_octant = hash_region(px, py, pz, tree_size * 0.5f);
    return true;
}

// Not part of the reference either: with all three scale factors NaN (a zero or NaN direction) every DDA distance is NaN,
// `min_step == distance` holds on no axis and nothing ever steps. One finite factor is enough (f32::min ignores NaN).
__device__ __forceinline__ bool no_usable_scale_factor(const RayConst& r) { return r.sfx != r.sfx && r.sfy != r.sfy && r.sfz != r.sfz; }

// root_entry followed by ray_setup, as the viewport kernels call them. (Sharing one refined reciprocal per divisor among
// the twelve divisions was built, validated bit for bit and measured in round 2: 0-1.5 % slower on every scene - the range
// guards cost what the shorter division sequences saved - and removed; profiles/r02_experiment_shared_rcp_far_plane.log.)
__device__ __forceinline__ bool root_entry_and_setup(RayConst& r, float tree_size, float& px, float& py, float& pz,
                                                     uint32_t& target_octant) {
    if (!root_entry(r, tree_size, px, py, pz, target_octant)) return false;
    ray_setup(r);
    return !no_usable_scale_factor(r);
}

// ---------------------------------------------------------------------------------------------------------------
// Exact fast-forward of the reference's "0.1 nudge" crawl (raytracing_on_cpu.rs:548-562).
//
// When the root's occupancy test fails (:438-443) the root is popped, the stack is empty, and the outer loop only does
//     p += direction * 0.1 ; bounds check ; re-hash the root octant
// once per iteration until the 4x4x4 bitmap cell of p changes (the test depends on nothing else) or p leaves the cube.
// A ray looking over a 1024^3 terrain repeats this ~10^4 times. The additions are f32 and must be reproduced bit for
// bit, but they have a closed form: for x in one binade [2^e, 2^(e+1)), ulp u = 2^(e-23), x = X*u with integer X, and
// c = t*u with real t, RN(x + c) = (X + rint(t))*u whenever t is not a tie (frac(t) != 0.5; ties are regular too once
// X is even, see crawl_limit) and the sum stays in the binade - independent of X. So n consecutive additions give (X + n*q)*u exactly, q = rint(t). crawl_limit() returns
// how many additions one axis can take before it would leave its binade or cross the next cell / cube boundary
// (multiples of size/4); the minimum over the axes is applied in one step. Anything irregular (a tie on an odd
// mantissa, denormals, |c| >= 2^e) returns 0 and the caller falls back to one explicit addition, which is always valid.
struct CrawlAxis {
    uint32_t limit;  // additions that are certainly "same binade, same cell, in bounds"
    int q;           // mantissa increment per addition
};

// floor(gap / q) from below without an integer division: the approximate quotient is at most 2 ulp high, the
// (1 - 2^-20) factor pulls it under the real quotient, and truncation then never exceeds floor(gap / q). A smaller
// count is always valid (fewer additions are fast-forwarded). gap < 2^24 and q < 2^23 are exact in f32.
__device__ __forceinline__ uint32_t floor_div_below(uint32_t gap, uint32_t q) {
    return (uint32_t)(__fdividef((float)gap, (float)q) * 0.99999904632568359375f);
}

__device__ __forceinline__ CrawlAxis crawl_limit(float x, float c, float quarter, float inv_quarter) {
    CrawlAxis a;
    a.limit = 0u;
    a.q = 0;
    const uint32_t bits = __float_as_uint(x);
    const uint32_t ef = bits >> 23;            // sign is 0: x > 0 inside the cube
    if (ef < 24u || ef > 200u) return a;       // denormal / tiny / huge: step explicitly
    const uint32_t X = (bits & 0x7FFFFFu) | 0x800000u;
    const float to_ulps = __uint_as_float((277u - ef) << 23);  // 2^(23 - e), exact
    const float t = c * to_ulps;               // exact scaling
    if (!(fabsf(t) < 8388608.0f)) return a;    // |c| >= 2^e (or NaN): leaves the binade at once
    // q = rint(t), ties to even. If t is an exact tie (frac = 0.5) the sum (X + t)*u rounds to the EVEN mantissa of
    // X + floor(t), X + ceil(t). With X even that is X + q (q is the even candidate), and X stays even afterwards, so
    // the recurrence is regular again; with X odd one explicit addition makes it even first.
    const int q = __float2int_rn(t);
    if (fabsf(t - (float)q) == 0.5f && (X & 1u)) return a;
    a.q = q;
    if (q == 0) {                              // x + c rounds back to x: this axis never moves
        a.limit = 0xFFFFFFFFu;
        return a;
    }
    const float cell = floorf(x * inv_quarter);  // bitmap cell along this axis (x * 4 / size, exact scaling)
    if (q > 0) {
        // stays in the binade while X_n <= 2^24 - 1, and in the cell (for the last cell: inside the cube) while
        // x_n < (cell + 1) * size/4. The boundary in ulps is an exact integer, possibly beyond the binade.
        const float bu = ((cell + 1.0f) * quarter) * to_ulps;
        const uint32_t top = bu >= 16777216.0f ? 16777216u : (uint32_t)bu;
        a.limit = (top - 1u >= X) ? floor_div_below(top - 1u - X, (uint32_t)q) : 0u;
    } else {
        // stays in the binade while X_n >= 2^23 and in the cell while x_n >= cell * size/4 (x_n > 0 is implied)
        const uint32_t b = (uint32_t)((cell * quarter) * to_ulps);
        const uint32_t bottom = max(b, 0x800000u);
        a.limit = (X >= bottom) ? floor_div_below(X - bottom, (uint32_t)(-q)) : 0u;
    }
    return a;
}

__device__ __forceinline__ float crawl_apply(float x, int q, uint32_t n) {
    if (q == 0) return x;
    const uint32_t bits = __float_as_uint(x);
    const uint32_t X = (bits & 0x7FFFFFu) | 0x800000u;
    const uint32_t Xn = (uint32_t)((int)X + q * (int)n);  // stays within [2^23, 2^24) by construction
    return __uint_as_float((bits & 0x7F800000u) | (Xn & 0x7FFFFFu));
}

// The level-of-detail test of get_by_ray_at_lod, raytracing_on_cpu.rs:370-376:
//     mip_level < (ray.origin - (p / (mip_level * 2.)).round() * (mip_level * 2.)).length() / viewing_distance
// mip_level * 2 is not a power of two in general: IEEE division, f32::round = roundf (half away from zero). A
// mip_level of 0 gives inf / NaN operands and a false comparison, a negative one (the level drifts, see below) just
// computes - both exactly as in the reference.
//
// Most evaluations are far from the boundary, so a bracketing pre-test decides them without the three IEEE divisions:
// a (p rounded to the grid of 2L) lies within sqrt(3) L of p - plus at most 2^-22 |p| per axis from the rounded quotient
// and product, < 1 for trees up to 2^20 - so D = |origin - a| is within s = 1.75 L + 1 of d = |origin - p|. With
// t = L vd the test "L < D / vd" is certainly false when t - s > d and certainly true when d - s > t; both sides carry
// a 1e-5 relative margin, two orders above what the f32 evaluation of either form can be off by. Everything in between,
// and every odd input (L < 1, vd <= 0 or NaN, huge trees), takes the exact form.
__device__ __forceinline__ bool lod_wants_mip(const RayConst& r, float px, float py, float pz, float mip_level,
                                              float viewing_distance, float tree_size) {
    if (mip_level >= 1.0f && viewing_distance > 0.0f && tree_size <= 1048576.0f) {
        const float vx = px - r.ox, vy = py - r.oy, vz = pz - r.oz;
        const float d2 = (vx * vx) + (vy * vy) + (vz * vz);
        const float t = mip_level * viewing_distance, s = 1.75f * mip_level + 1.0f;
        const float u = t * 0.99999f - s;
        if (u > 0.0f && u * u > d2 * 1.00001f) return false;
        const float v = t * 1.00001f + s;
        if (d2 * 0.99999f > v * v) return true;
    }
    const float m2 = mip_level * 2.0f;
    const float ax = roundf(px / m2) * m2, ay = roundf(py / m2) * m2, az = roundf(pz / m2) * m2;
    const float wx = r.ox - ax, wy = r.oy - ay, wz = r.oz - az;
    return mip_level < sqrtf((wx * wx) + (wy * wy) + (wz * wz)) / viewing_distance;
}

// "The LOD test stays false for the rest of this crawl": lets the LOD variant use the crawl fast-forward.
// While the root keeps failing its occupancy test, every iteration adds 1 to mip_level (L) and direction * 0.1 to p.
// The test is L' < |origin - a'| / vd with a' = p' rounded to the grid of 2L', so |p' - a'| <= sqrt(3) L', and p' stays
// inside the cube, so |origin - p'| <= |origin - p| + sqrt(3) size. It is therefore false whenever
//     L' (1 - sqrt(3) / vd) >= (|origin - p| + sqrt(3) size) / vd ,
// which for vd >= 4 follows from L' >= L >= 1.77 (|origin - p| + 1.74 size) / vd. The check below asks for
// L vd >= 4 (|origin - p| + 2 size): more than twice that, far beyond the few ulp the f32 evaluation of the test can be
// off by. L < 4e6 keeps L + n an exact integer in f32 for any fast-forward count n (< 2^23).
__device__ __forceinline__ bool lod_quiescent(const RayConst& r, float px, float py, float pz, float mip_level,
                                              float viewing_distance, float tree_size) {
    if (!(viewing_distance >= 4.0f) || !(mip_level < 4.0e6f)) return false;
    const float wx = px - r.ox, wy = py - r.oy, wz = pz - r.oz;
    const float reach = sqrtf((wx * wx) + (wy * wy) + (wz * wz)) + 2.0f * tree_size;
    return mip_level * viewing_distance >= 4.0f * reach;  // NaN / 0 * inf compare false
}

// The loops of get_by_ray_at_lod, raytracing_on_cpu.rs:349-565, entered with the point / octant root_entry produced
// and a fully set-up RayConst.
//   LOD = false: the tree's MIP maps are disabled, so the level-of-detail branch (:368-386) is dead code and
//                viewing_distance is irrelevant - Octree::get_by_ray on a tree as every BASELINE config builds it.
//   LOD = true : MIP maps enabled. `mip_level` follows the reference literally: it starts at log2(size / dim) (:349),
//                loses 1 per PUSH below the root and gains 1 per POP including the root's, and is NOT reset when the
//                walk restarts from the root - so it drifts upwards by one per root cycle and downwards whenever the
//                4-entry ring stack has dropped entries. Every failing root iteration of the crawl also raises
//                mip_level and re-evaluates the LOD test, so the crawl fast-forward runs only once lod_quiescent()
//                proves the test false for the rest of the crawl (then n nudges are also n increments of mip_level).
template <bool LOD, int BS = -1>
__device__ __forceinline__ bool traverse(const DeviceTree& t, const RayConst& r, float px, float py, float pz,
                                         uint32_t target_octant, TraceResult& out, float viewing_distance = 0.0f) {
    const float tree_size = (float)t.tree_size;
    // log2 of a power of two: (size / dim) = 2^k exactly, k read from the exponent field
    float mip_level = LOD ? (float)((int)((__float_as_uint(tree_size * inv_brick_dim_of<BS>(t)) >> 23) & 0xFFu) - 127) : 0.0f;
    // NodeStack<u32, 4>, raytracing_on_cpu.rs:20-82: a ring buffer that overwrites its oldest entry. Pushes are always a
    // child of the current top, so the valid entries are the current node and its nearest ancestors, at most four; what
    // the ring loses when it wraps is only HOW MANY of them can still be popped. The kernel keeps that count and takes
    // the entry below the top from the node record (NodeHead::aux = parent index) - the same nodes in the same order,
    // including the early "stack empty" restarts from the root in trees deeper than four levels.
    uint32_t count = 0;
    uint32_t cur = 0;
    float bx = 0.0f, by = 0.0f, bz = 0.0f, bsize = tree_size;
    float binv = t.inv_tree_size;  // 1 / bsize, exact (powers of two), tracked alongside bsize

    // the root record is needed on every restart; Internal / Nothing roots can crawl (a leaf root probes bricks)
    const uint4 root_hd = node_head_of(t.node_rec);
    const bool root_can_crawl = (root_hd.z & 3u) == NK_INTERNAL || (root_hd.z & 3u) == NK_NOTHING;
    const float cwx = r.dx * 0.1f, cwy = r.dy * 0.1f, cwz = r.dz * 0.1f;  // `ray.direction * 0.1` (:551)
    const float quarter = tree_size * 0.25f, inv_quarter = t.inv_tree_size * 4.0f;
    // Defensive cap on restarts from the root (not in the reference): once `direction * 0.1` is below half an ulp of p on
    // every axis (trees of 2^20 and more) the nudge no longer moves p and the reference's outer loop never ends. 2^26
    // iterations are far beyond any ray that makes progress (a 2^19 tree is crossed in < 2^24 nudges); then: a miss.
    uint32_t restart_budget = 1u << 26;

    while (target_octant != OOB_OCTANT) {
        if (root_can_crawl) {
            // Outer iterations in which the root is popped right away, without touching the stack or the tree:
            // test the root's occupancy for the cell of p; on failure nudge p, check the bounds, repeat.
            // Per-axis fast-forward state: q = mantissa increment per nudge, rem = further nudges that provably keep the
            // axis in its binade, bitmap cell and the cube. An axis is re-analysed only after its own budget ran out.
            uint32_t fails = 0u;
            int qx = 0, qy = 0, qz = 0;
            uint32_t remx = 0u, remy = 0u, remz = 0u;
            for (;;) {
                if (--restart_budget == 0u) {
                    out.palette_value = NIL;
                    return false;
                }
                // `p * 4 / size`: two scalings by powers of two = one by their (exact) product, inv_quarter
                const float cpx = rust_clamp(px * inv_quarter, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
                const float cpy = rust_clamp(py * inv_quarter, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
                const float cpz = rust_clamp(pz * inv_quarter, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
                // LOD: the root's MIP is probed before the occupancy test (:368-386) - leave that to the node loop
                if (LOD && lod_wants_mip(r, px, py, pz, mip_level, viewing_distance, tree_size)) break;
                if ((root_hd.x | root_hd.y) != 0u &&
                    ray_may_hit_node(t, root_hd.x, root_hd.y, bitmap_coord_of_clamped(cpx), bitmap_coord_of_clamped(cpy), bitmap_coord_of_clamped(cpz), r.dirbits))
                    break;  // the root survives its test: run the node loop below
                bool regular = true;
                if (LOD) {
                    mip_level += 1.0f;  // the root's POP (:447)
                    regular = lod_quiescent(r, px, py, pz, mip_level, viewing_distance, tree_size);
                }
                if (regular && ++fails >= 2u) {
                    // a run of failing iterations: apply as many nudges as provably change nothing, at once
                    if (remx == 0u) { const CrawlAxis a = crawl_limit(px, cwx, quarter, inv_quarter); qx = a.q; remx = a.limit; }
                    if (remy == 0u) { const CrawlAxis a = crawl_limit(py, cwy, quarter, inv_quarter); qy = a.q; remy = a.limit; }
                    if (remz == 0u) { const CrawlAxis a = crawl_limit(pz, cwz, quarter, inv_quarter); qz = a.q; remz = a.limit; }
                    const uint32_t n = min(min(remx, remy), remz);
                    if (n == 0xFFFFFFFFu) {  // the nudge rounds away on every axis: p can never change again (the reference spins)
                        out.palette_value = NIL;
                        return false;
                    }
                    if (n != 0u) {
                        px = crawl_apply(px, qx, n);
                        py = crawl_apply(py, qy, n);
                        pz = crawl_apply(pz, qz, n);
                        if (LOD) mip_level += (float)n;  // n more root POPs; exact: integers below 2^24
                        if (remx != 0xFFFFFFFFu) remx -= n;
                        if (remy != 0xFFFFFFFFu) remy -= n;
                        if (remz != 0xFFFFFFFFu) remz -= n;
                    }
                    // the explicit nudge below consumes one more regular step of every axis that still has budget
                    if (remx != 0u && remx != 0xFFFFFFFFu) remx -= 1u;
                    if (remy != 0u && remy != 0xFFFFFFFFu) remy -= 1u;
                    if (remz != 0u && remz != 0xFFFFFFFFu) remz -= 1u;
                }
                px = px + cwx;
                py = py + cwy;
                pz = pz + cwz;
                if (!(px < tree_size && py < tree_size && pz < tree_size && px > 0.0f && py > 0.0f && pz > 0.0f)) {
                    out.palette_value = NIL;
                    return false;
                }
            }
            target_octant = hash_region(px, py, pz, tree_size * 0.5f);
        }
        cur = 0;
        bx = by = bz = 0.0f;
        bsize = tree_size;
        binv = t.inv_tree_size;
        count = min(count + 1u, 4u);  // node_stack.push(root)
        while (count != 0u) {
            // cur == top of the stack here: one 16-byte load serves occupancy bits and node kind
            const uint4* rec = node_record(t, cur);
            const uint4 hd = node_head_of(rec);
            const uint32_t oc_lo = hd.x, oc_hi = hd.y, meta = hd.z;
            const uint32_t kind = meta & 3u;
            if (LOD) {
                // :368-386 far enough away, the node's MIP brick stands in for its content. A miss leaves the point
                // where the brick walk ended and target_octant as it was.
                if (lod_wants_mip(r, px, py, pz, mip_level, viewing_distance, tree_size)) {
                    const uint32_t mkind = (meta >> 18) & 3u;
                    if (mkind != BK_EMPTY &&
                        probe_brick<BS>(t, r, px, py, pz, mkind, __ldg(t.node_mip + cur), bx, by, bz, bsize, binv, out))
                        return true;
                }
            }
#if SVX_SINGLE_PROBE_SITE
            if (target_octant != OOB_OCTANT && kind >= NK_LEAF) {
                // One call site for both leaf kinds (:393-421): the brick walk is instantiated once, and lanes that sit in a
                // UniformLeaf walk their bricks together with lanes that sit in an octant of a Leaf.
                const bool uniform = kind == NK_UNIFORM;
                const uint32_t o = uniform ? 0u : target_octant;
                const uint32_t bkind = (meta >> (2u + 2u * o)) & 3u;
                if (bkind != BK_EMPTY) {
                    // child_bounds_for: min + offset * size / 2 with offset 0 or 1 per axis = min or min + size/2
                    const float hs = bsize * 0.5f;
                    const float cbx = (!uniform && (o & 1u)) ? bx + hs : bx;
                    const float cby = (!uniform && (o & 4u)) ? by + hs : by;
                    const float cbz = (!uniform && (o & 2u)) ? bz + hs : bz;
                    if (probe_brick<BS>(t, r, px, py, pz, bkind, node_slot_of(rec, o), cbx, cby, cbz, uniform ? bsize : hs,
                                        uniform ? binv : binv * 2.0f, out))
                        return true;
                }
            }
#else
            if (target_octant != OOB_OCTANT) {
                if (kind == NK_UNIFORM) {
                    const uint32_t ubk = (meta >> 2) & 3u;
                    if (ubk != BK_EMPTY && probe_brick<BS>(t, r, px, py, pz, ubk, node_slot_of(rec, 0u), bx, by, bz, bsize, binv, out)) return true;
                } else if (kind == NK_LEAF) {
                    const uint32_t bkind = (meta >> (2u + 2u * target_octant)) & 3u;
                    if (bkind != BK_EMPTY) {
                        const float hs = bsize * 0.5f;
                        const uint32_t slot = node_slot_of(rec, target_octant);
                        // child_bounds_for: min + offset * size / 2 with offset 0 or 1 per axis = min or min + size/2
                        if (probe_brick<BS>(t, r, px, py, pz, bkind, slot, (target_octant & 1u) ? bx + hs : bx,
                                        (target_octant & 4u) ? by + hs : by, (target_octant & 2u) ? bz + hs : bz, hs,
                                        binv * 2.0f, out))
                            return true;
                    }
                }
            }
#endif
            // position inside the node in 4x4x4 bitmap cells (:425-436)
            // `(p - min) * 4 / size`: the two scalings by powers of two are one by their exact product 4 / size
            const float cells = 4.0f * binv;
#if SVX_PACKED_DDA
            float bpx, bpy;
            unpack2(mul2(sub2(pack2(px, py), pack2(bx, by)), pack2(cells, cells)), bpx, bpy);
            bpx = rust_clamp(bpx, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
            bpy = rust_clamp(bpy, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
#else
            float bpx = rust_clamp((px - bx) * cells, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
            float bpy = rust_clamp((py - by) * cells, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
#endif
            float bpz = rust_clamp((pz - bz) * cells, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE);
            if (kind == NK_UNIFORM || target_octant == OOB_OCTANT || (oc_lo | oc_hi) == 0u ||
                !ray_may_hit_node(t, oc_lo, oc_hi, bitmap_coord_of_clamped(bpx), bitmap_coord_of_clamped(bpy), bitmap_coord_of_clamped(bpz), r.dirbits)) {
                // POP (:445-474)
                if (LOD) mip_level += 1.0f;
                count -= 1u;
                if (count != 0u) {
                    cur = hd.w;  // the entry below the top of the stack: this node's parent
                    // parent bounds: min - min % (2*size), size * 2 (:452-456, :470-471) - exact integer-valued f32, i.e.
                    // the parent node's own bounds, which the serialiser stored (node_bounds). The octant the node
                    // occupies in its parent, hash_region(centre - parent min, size) (:458-463), is a property of the tree:
                    // the serialiser stored it in the node's meta word.
                    const float4 pb = node_bounds_of(node_record(t, cur));
                    const uint32_t from = (meta >> 20) & 7u;
                    bool sx, sy, sz;
                    dda_step(r, px, py, pz, bx, by, bz, bsize, sx, sy, sz);
                    target_octant = step_octant(from, sx, sy, sz, r.dirbits >> 3);
                    bsize = pb.w;
                    binv = binv * 0.5f;
                    bx = pb.x; by = pb.y; bz = pb.z;
                }
                continue;
            }
            const float hs = bsize * 0.5f;
            float tbx = (target_octant & 1u) ? bx + hs : bx;
            float tby = (target_octant & 4u) ? by + hs : by;
            float tbz = (target_octant & 2u) ? bz + hs : bz;
            // NodeChildren::child(): only Internal nodes carry child keys (node.rs:49-54)
            uint32_t child = (kind == NK_INTERNAL) ? node_slot_of(rec, target_octant) : NIL;
            if (child != NIL && octant_occupied(oc_lo, oc_hi, target_octant)) {
                // PUSH (:484-492)
                cur = child;
                bx = tbx; by = tby; bz = tbz;
                bsize = hs;
                binv = binv * 2.0f;
#if SVX_PACKED_DDA
                float rx, ry;
                unpack2(sub2(pack2(px, py), pack2(bx, by)), rx, ry);
                target_octant = hash_region(rx, ry, pz - bz, hs * 0.5f);
#else
                target_octant = hash_region(px - bx, py - by, pz - bz, hs * 0.5f);
#endif
                count = min(count + 1u, 4u);  // node_stack.push(child)
                if (LOD) mip_level -= 1.0f;
            } else {
                // ADVANCE (:497-544)
                // `step * 4. / size` is +-cells or +0 (sic: 4/size cells)
                float qx = with_sign_of(cells, r.dx), qy = with_sign_of(cells, r.dy), qz = with_sign_of(cells, r.dz);
#ifndef SVX_HOST_MIRROR
                asm volatile("" : "+f"(qx), "+f"(qy), "+f"(qz));  // loop constants in registers, not rebuilt per step
#endif
                // child_bounds_for(target_octant) (:506) moves by exactly +-size/2 along every stepped axis (integers: exact)
                const float hx = with_sign_of(hs, r.dx), hy = with_sign_of(hs, r.dy), hz = with_sign_of(hs, r.dz);
#if SVX_PACKED_DDA
                uint64_t offxy = pack2(fmaxf(hx, 0.0f), fmaxf(hy, 0.0f));  // negative ? 0 : hs
                float offz = fmaxf(hz, 0.0f);
#ifndef SVX_HOST_MIRROR
                asm volatile("" : "+l"(offxy), "+f"(offz));  // loop constants in registers, not rebuilt per step
#endif
                uint64_t tbxy = pack2(tbx, tby);  // carried as a pair: the sibling's bounds are not needed after the walk
#endif
                for (;;) {
                    bool sx, sy, sz;
#if SVX_PACKED_DDA
                    const DdaStep st = dda_step_off(r, px, py, pz, tbxy, tbz, offxy, offz, sx, sy, sz);
                    target_octant = step_octant(target_octant, sx, sy, sz, r.dirbits >> 3);
                    if (target_octant == OOB_OCTANT) break;
                    unpack2(tbxy, tbx, tby);
                    add_both_if_equal(st.m, st.d_x, tbx, hx, bpx, qx);
                    add_both_if_equal(st.m, st.d_y, tby, hy, bpy, qy);
                    add_both_if_equal(st.m, st.d_z, tbz, hz, bpz, qz);
                    tbxy = pack2(tbx, tby);
#else
                    dda_step(r, px, py, pz, tbx, tby, tbz, hs, sx, sy, sz);
                    target_octant = step_octant(target_octant, sx, sy, sz, r.dirbits >> 3);
                    if (target_octant == OOB_OCTANT) break;
                    if (sx) { tbx = tbx + hx; bpx = bpx + qx; }
                    if (sy) { tby = tby + hy; bpy = bpy + qy; }
                    if (sz) { tbz = tbz + hz; bpz = bpz + qz; }
#endif
                    if (kind == NK_INTERNAL) {
                        child = node_slot_of(rec, target_octant);
                        if (child != NIL && octant_occupied(oc_lo, oc_hi, target_octant) &&
                            ray_may_hit_node(t, oc_lo, oc_hi, bitmap_coord(bpx), bitmap_coord(bpy), bitmap_coord(bpz), r.dirbits))
                            break;
                    } else if (kind == NK_LEAF) {
                        if (((meta >> (2u + 2u * target_octant)) & 3u) != BK_EMPTY) break;
                    }
                }
            }
        }
        // restart from the root after a 0.1 nudge (:548-562)
        if (--restart_budget == 0u) break;
        px = px + cwx;
        py = py + cwy;
        pz = pz + cwz;
        if (px < tree_size && py < tree_size && pz < tree_size && px > 0.0f && py > 0.0f && pz > 0.0f)
            target_octant = hash_region(px, py, pz, tree_size * 0.5f);
        else
            target_octant = OOB_OCTANT;
    }
    out.palette_value = NIL;
    return false;
}

// Octree::get_by_ray / get_by_ray_at_lod (raytracing_on_cpu.rs:316-325) for a ray whose origin / direction are set in `r`
template <bool LOD>
__device__ __forceinline__ bool trace_ray(const DeviceTree& t, RayConst& r, TraceResult& out, float viewing_distance = 0.0f) {
    float px, py, pz;
    uint32_t target_octant;
    out.palette_value = NIL;
    if (certain_root_miss(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, (float)t.tree_size)) return false;
    if (!root_entry(r, (float)t.tree_size, px, py, pz, target_octant)) return false;
    ray_setup(r);
    if (no_usable_scale_factor(r)) return false;
    return traverse<LOD>(t, r, px, py, pz, target_octant, out, viewing_distance);
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
