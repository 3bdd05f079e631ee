// TEST INFRASTRUCTURE, never part of libshocovox_b200: the kernels' device code (shocovox_b200/csrc/traverse.cuh, the
// very header the CUDA kernels are built from) compiled by the HOST compiler, so that its logic - the transformed DDA
// arithmetic, the mirrored brick walk, the parent-index node stack, the crawl fast-forward, the brick-dimension
// instantiations - can be checked against the oracle in the CPU test suite, without a GPU. What it cannot show is what
// nvcc / ptxas make of the same source (FMA contraction, packed instructions): that stays with the `-m gpu` parity tests.
// The product has no CPU ray path; nothing outside tests/ builds or loads this file.
//
// Mirrors rays_body<LOD> of csrc/kernels.cu (Octree::get_by_ray / get_by_ray_at_lod, reference
// src/raytracing/raytracing_on_cpu.rs:316-565).
#define SVX_HOST_MIRROR 1
#include <math.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include <cuda_runtime.h>  // vector types; __device__ / __forceinline__ expand to host-compiler attributes

// ---- the CUDA intrinsics traverse.cuh uses, with the semantics of the PTX instructions they stand for -------------------
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __float_as_uint(float v) { uint32_t u; std::memcpy(&u, &v, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float v; std::memcpy(&v, &u, 4); return v; }
// cvt.{rzi,rmi,rni}.s32.f32: saturating, NaN -> 0
static inline int svx_saturate_to_int(float rounded) {
    if (rounded != rounded) return 0;
    if (rounded >= 2147483648.0f) return 2147483647;
    if (rounded <= -2147483648.0f) return -2147483647 - 1;
    return (int)rounded;
}
static inline int __float2int_rz(float v) { return svx_saturate_to_int(truncf(v)); }
static inline int __float2int_rd(float v) { return svx_saturate_to_int(floorf(v)); }
static inline int __float2int_rn(float v) { return svx_saturate_to_int(nearbyintf(v)); }  // default rounding mode: ties to even
static inline float __fdividef(float a, float b) { return a / b; }  // the kernels only use it where any close quotient is valid
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
template <typename T>
static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T>
static inline T max(T a, T b) { return a < b ? b : a; }

#include "kernels.cuh"   // RayHitRecord
#include "traverse.cuh"
#ifdef SVX_MIRROR_RESUMABLE  // = node-loop iterations per call: the suspend/resume form the lane-refill schedule runs
#include "traverse_refill.cuh"
#endif

namespace {

template <bool LOD, int BS>
void trace_range(const svx::DeviceTree& tree, const float* rays, uint64_t begin, uint64_t end, float viewing_distance,
                 svx::RayHitRecord* out) {
    using namespace svx;
    for (uint64_t i = begin; i < end; ++i) {
        RayConst r;
        r.ox = rays[6 * i + 0]; r.oy = rays[6 * i + 1]; r.oz = rays[6 * i + 2];
        r.dx = rays[6 * i + 3]; r.dy = rays[6 * i + 4]; r.dz = rays[6 * i + 5];
        TraceResult res;
        res.palette_value = NIL;
        bool hit = false;
        float px, py, pz;
        uint32_t target_octant;
        // trace_ray (traverse.cuh) with the brick dimension as a template argument, as the viewport kernels instantiate it
        if (!certain_root_miss(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, (float)tree.tree_size) &&
            root_entry_and_setup(r, (float)tree.tree_size, px, py, pz, target_octant)) {
#ifdef SVX_MIRROR_RESUMABLE
            TraverseState S;
            traverse_begin<LOD, BS>(tree, S, px, py, pz, target_octant);
            int walk;
            do walk = traverse_resumable<LOD, BS>(tree, r, S, res, viewing_distance, (uint32_t)(SVX_MIRROR_RESUMABLE));
            while (walk == WALK_SUSPENDED);
            hit = walk == WALK_HIT;
#else
            hit = traverse<LOD, BS>(tree, r, px, py, pz, target_octant, res, viewing_distance);
#endif
        }
        RayHitRecord h;
        h.hit = hit ? 1u : 0u;
        h.palette_value = hit ? res.palette_value : NIL;
        h.impact[0] = h.impact[1] = h.impact[2] = 0.0f;
        h.normal[0] = h.normal[1] = h.normal[2] = 0.0f;
        h.distance = 0.0f;
        if (hit) {
            h.impact[0] = res.px; h.impact[1] = res.py; h.impact[2] = res.pz;
            impact_normal(res, h.normal[0], h.normal[1], h.normal[2]);
            const float vx = res.px - r.ox, vy = res.py - r.oy, vz = res.pz - r.oz;
            h.distance = sqrtf((vx * vx) + (vy * vy) + (vz * vz));
        }
        out[i] = h;
    }
}

}  // namespace

// specialise != 0: use the compile-time brick dimension instantiation when the tree's dimension is 8 or 32 (what
// launch_render does); 0: always the generic code. node_mip != NULL: the tree's MIP maps are enabled - the LOD
// instantiation runs, as in the library (DeviceTree::mips_enabled selects it, not the viewing distance).
// Returns 0, or 1 for inconsistent arguments.
extern "C" int svx_host_mirror_get_by_rays(const uint32_t* node_rec, uint32_t n_nodes, const uint32_t* voxels, const uint32_t* bits,
                                           uint32_t n_bricks, const uint32_t* ray_lut, uint32_t tree_size, uint32_t brick_dim,
                                           int specialise, const float* rays, uint64_t n, svx::RayHitRecord* out, int threads,
                                           const uint32_t* node_mip, float viewing_distance) {
    if (!node_rec || !ray_lut || !rays || !out || n_nodes == 0 || brick_dim == 0 || (brick_dim & (brick_dim - 1))) return 1;
    svx::DeviceTree t{};
    t.node_rec = reinterpret_cast<const uint4*>(node_rec);
    t.node_mip = node_mip;
    t.voxels = voxels;
    t.brick_bits = bits;
    t.palette = nullptr;
    t.ray_lut = reinterpret_cast<const uint2*>(ray_lut);
    t.n_nodes = n_nodes;
    t.n_bricks = n_bricks;
    t.tree_size = tree_size;
    t.brick_dim = brick_dim;
    t.brick_shift = 0;
    while ((1u << t.brick_shift) < brick_dim) ++t.brick_shift;
    t.brick_dim_sq = brick_dim * brick_dim;
    t.bit_words = (brick_dim * brick_dim * brick_dim + 31u) / 32u;
    t.n_colors = 0;
    t.mips_enabled = node_mip ? 1u : 0u;
    t.inv_tree_size = 1.0f / (float)tree_size;
    t.inv_brick_dim = 1.0f / (float)brick_dim;
    const int nt = std::max(1, threads);
    std::vector<std::thread> pool;
    for (int k = 0; k < nt; ++k) {
        const uint64_t b = n * k / nt, e = n * (k + 1) / nt;
        pool.emplace_back([=] {
            const uint32_t shift = specialise ? t.brick_shift : 0xFFFFFFFFu;
            if (node_mip) {
                if (shift == 3u) trace_range<true, 3>(t, rays, b, e, viewing_distance, out);
                else if (shift == 5u) trace_range<true, 5>(t, rays, b, e, viewing_distance, out);
                else trace_range<true, -1>(t, rays, b, e, viewing_distance, out);
            } else {
                if (shift == 3u) trace_range<false, 3>(t, rays, b, e, viewing_distance, out);
                else if (shift == 5u) trace_range<false, 5>(t, rays, b, e, viewing_distance, out);
                else trace_range<false, -1>(t, rays, b, e, viewing_distance, out);
            }
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}
