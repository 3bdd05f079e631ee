// sm_100a kernels of the primary-ray path: the viewport kernel (ray generation + traversal + framebuffer) and the
// batched get_by_ray kernel. Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (see build.py).
#include <algorithm>

#include "kernels.cuh"
#include "traverse.cuh"
#include "traverse_refill.cuh"

namespace svx {

// One warp renders an 8x4 pixel tile (coherent rays, 32-byte framebuffer segments per row and plane).
// CTA shape: SVX_BLOCK_WARPS_X x SVX_BLOCK_WARPS_Y warps, each warp an 8x4 pixel tile. Small CTAs retire as soon as
// their slowest warp is done, which matters because neighbouring rays can differ a lot in length (measured on B200:
// 128-thread CTAs of 2x2 or 4x1 warps beat 4x2 and 4x4 by 2-5 % on every scene; 64- and 32-thread CTAs lose again).
#ifndef SVX_BLOCK_WARPS_X
#define SVX_BLOCK_WARPS_X 2
#endif
#ifndef SVX_BLOCK_WARPS_Y
#define SVX_BLOCK_WARPS_Y 2
#endif
#ifndef SVX_MIN_BLOCKS
#define SVX_MIN_BLOCKS (32 / (SVX_BLOCK_WARPS_X * SVX_BLOCK_WARPS_Y))  // 1024 threads (64 registers each) per SM
#endif
#ifndef SVX_TICKET_TILES
#define SVX_TICKET_TILES 1u   // tiles per ticket of the persistent schedule (1, 2, 4 or 8); measured: 1 is best, larger tickets lengthen the tail
#endif
// Shape of the pixel tile one warp renders: 2^SVX_WARP_LOG2_W x 2^(5 - SVX_WARP_LOG2_W) pixels (3: 8x4, 4: 16x2, 5: 32x1)
#ifndef SVX_WARP_LOG2_W
#define SVX_WARP_LOG2_W 3
#endif
constexpr int WARP_LW = SVX_WARP_LOG2_W, WARP_LH = 5 - SVX_WARP_LOG2_W;
constexpr int WARP_W = 1 << WARP_LW, WARP_H = 1 << WARP_LH;
constexpr int TILE_W = WARP_W * SVX_BLOCK_WARPS_X;
constexpr int TILE_H = WARP_H * SVX_BLOCK_WARPS_Y;
constexpr int BLOCK_THREADS = 32 * SVX_BLOCK_WARPS_X * SVX_BLOCK_WARPS_Y;

__device__ __forceinline__ void pixel_of_thread(int& tx, int& ty) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tx = ((warp % SVX_BLOCK_WARPS_X) << WARP_LW) + (lane & (WARP_W - 1));
    ty = ((warp / SVX_BLOCK_WARPS_X) << WARP_LH) + (lane >> WARP_LW);
}

// Ray generation of the caller loop, reference examples/cpu_render.rs:104-114, in its f32 operation order:
//   glass_point = (bottom_left + (right * x as f32) * pixel_width) + (up * y as f32) * pixel_height
//   direction   = (glass_point - origin).normalized()          (three divisions, vector.rs:79-81)
// Returns the un-normalised vector glass_point - origin; the exact normalisation happens only for rays that may hit.
__device__ __forceinline__ void glass_vector(const FrameParams& f, uint32_t x, uint32_t y, float& vx, float& vy, float& vz) {
    const float xf = (float)x, yf = (float)y;
    const float gx = (f.blx + (f.rx * xf) * f.pixel_width) + (f.ux * yf) * f.pixel_height;
    const float gy = (f.bly + (f.ry * xf) * f.pixel_width) + (f.uy * yf) * f.pixel_height;
    const float gz = (f.blz + (f.rz * xf) * f.pixel_width) + (f.uz * yf) * f.pixel_height;
    vx = gx - f.ox;
    vy = gy - f.oy;
    vz = gz - f.oz;
}

// SVX_PIXEL_TIMING=1 builds a MEASUREMENT variant of the library (tools/pixel_timing.py; never the product): every pixel's
// planes carry when its ray ran instead of what it hit - hit_id = SM cycles spent, albedo / distance = start / end of the
// pixel on the GPU's nanosecond timer (low 32 bits). The end is the WARP's: lanes that leave the traversal early are parked
// until the slowest ray of their warp is done (a stamp placed on the exit paths is not executed any earlier either - the
// scheduler runs the lanes still in the loop first), so these are warp times; what lanes lose inside a warp is in ncu's
// thread-efficiency counter.
#ifndef SVX_PIXEL_TIMING
#define SVX_PIXEL_TIMING 0
#endif
#if SVX_PIXEL_TIMING
__device__ __forceinline__ uint32_t timer_ns_lo() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (uint32_t)t;
}
#endif

// One pixel of the frame: ray generation, rejection tests, get_by_ray, framebuffer stores.
// `f32 as u8` of the caller loop's colour channels: truncating, saturating, NaN -> 0
__device__ __forceinline__ uint32_t channel_u8(float v) { return min(__float2uint_rz(v), 255u); }

// (x, lr) = column and shard-local row. LOD: the tree's MIP maps are enabled (traverse.cuh: traverse<LOD>).
// SHADE: also write the caller loop's shaded pixel (examples/cpu_render.rs:119-136) to the fourth plane.
// BS >= 0: the tree's brick dimension is the compile-time constant 2^BS (traverse.cuh: brick_dim_of); -1 = read it from the tree.
// shard-local row -> image row: the rank's k-th band is global band k * world + rank (plain interleave) or, rotated,
// k * world + (rank - k) mod world (FrameParams::band_rotate)
__device__ __forceinline__ uint32_t image_row_of(const FrameParams& f, uint32_t lr) {
    if (f.world == 1u) return lr;
    const uint32_t band = lr >> f.band_shift, within = lr & ((1u << f.band_shift) - 1u);
    uint32_t place = f.rank;
    if (f.band_rotate) {  // (rank - band) mod world without a division: the host checked that the frame has fewer than 2^16 bands
        const uint32_t t = f.rank + f.world - (band - __umulhi(band, f.world_magic) * f.world);
        place = t >= f.world ? t - f.world : t;
    }
    return ((band * f.world + place) << f.band_shift) + within;
}

// STAGED: the three values go to the CTA's staging area in shared memory (`stage[plane * 128 + slot]`) instead of the
// framebuffer; the caller writes them out as whole rows (render_staged_body). The other instantiations are unchanged by it.
// SHARDED = false: the launch renders a whole frame (world == 1), local rows are image rows and no band arithmetic is
// compiled in (the whole-frame kernels are sensitive to every instruction of this prologue: +1.3 % on minecraft 4K with the
// rotated interleave's code merely present behind a branch).
template <bool LOD, bool SHADE, int BS = -1, bool STAGED = false, bool SHARDED = true>
__device__ __forceinline__ void shade_pixel(const DeviceTree& tree, const FrameParams& f, uint32_t x, uint32_t lr,
                                            uint32_t* stage = nullptr, uint32_t slot = 0) {
    if (x >= f.width || lr >= f.rows_local) return;
    // shard-local row -> image row (interleaved bands of 2^band_shift rows); one GPU owns every row in order
    const uint32_t row = SHARDED ? image_row_of(f, lr) : lr;
    if (row >= f.height) return;
    const uint32_t y = f.height - 1u - row;  // pixel (x, y) lands in image row h-1-y (cpu_render.rs:106)
    const uint32_t i = (f.compact ? lr : row) * f.width + x;  // width * height < 2^32 (checked by the host)

#if SVX_PIXEL_TIMING
    const uint32_t timing_t0 = timer_ns_lo();
    const long long timing_c0 = clock64();
#endif
    uint32_t hit_id = NIL, rgba = 0u;
    float dist = 0.0f;
    // 0) pixels outside the projected bounding rectangle of the root cube are sky (host-computed, conservative)
    if (x < f.cull_x0 || x > f.cull_x1 || row < f.cull_row0 || row > f.cull_row1) {
        if constexpr (STAGED) {
            stage[slot] = NIL;
            stage[128u + slot] = 0u;
            stage[256u + slot] = 0u;
        } else {
            f.hit_id[i] = NIL;
            if (f.albedo) f.albedo[i] = 0u;
            f.distance[i] = 0.0f;
            if (SHADE) f.shaded[i] = 0xFF808080u;
        }
        return;
    }
    uint32_t pixel = 0xFF808080u;  // Rgb([128, 128, 128]) on a miss (cpu_render.rs:134)
    float vx, vy, vz;
    glass_vector(f, x, y, vx, vy, vz);
    const float tree_size = (float)tree.tree_size;
    // 1) cheap conservative rejection with an approximately normalised direction (see certain_root_miss); skipped when
    //    the host could bound the cube on the screen - inside that rectangle nearly every ray enters the cube anyway
    bool may_hit = true;
    if (f.prefilter) {
        const float rl = rsqrtf((vx * vx) + (vy * vy) + (vz * vz));
        may_hit = !certain_root_miss(f.ox, f.oy, f.oz, vx * rl, vy * rl, vz * rl, tree_size);
    }
    if (may_hit) {
        // 2) the reference's exact arithmetic for everything that may hit
        RayConst r;
        const float len = sqrtf((vx * vx) + (vy * vy) + (vz * vz));
        r.ox = f.ox; r.oy = f.oy; r.oz = f.oz;
        r.dx = vx / len; r.dy = vy / len; r.dz = vz / len;
        float px, py, pz;
        uint32_t target_octant;
        if (root_entry_and_setup(r, tree_size, px, py, pz, target_octant)) {
            TraceResult res;
            if (traverse<LOD, BS>(tree, r, px, py, pz, target_octant, res, f.viewing_distance)) {
                hit_id = res.palette_value;
                const uint32_t ci = res.palette_value & 0xFFFFu;
                if (ci < 0xFFFFu && ci < tree.n_colors) rgba = __ldg(tree.palette + ci);
                const float wx = res.px - r.ox, wy = res.py - r.oy, wz = res.pz - r.oz;
                dist = sqrtf((wx * wx) + (wy * wy) + (wz * wz));  // V3c::length, vector.rs:75-77
                if (SHADE) {
                    // diffuse_light_strength = 1. - (normal.dot(&light) / 2. + 0.5) ; channel = (c as f32 * strength) as u8
                    pixel = 0xFF000000u;  // a hit without a colour: the reference panics on albedo().unwrap()
                    if (ci < 0xFFFFu && ci < tree.n_colors) {
                        float nx, ny, nz;
                        impact_normal(res, nx, ny, nz);
                        const float strength = 1.0f - (((nx * f.lx + ny * f.ly) + nz * f.lz) * 0.5f + 0.5f);
                        pixel |= channel_u8((float)(rgba & 0xFFu) * strength) | (channel_u8((float)((rgba >> 8) & 0xFFu) * strength) << 8) |
                                 (channel_u8((float)((rgba >> 16) & 0xFFu) * strength) << 16);
                    }
                }
            }
        }
    }
#if SVX_PIXEL_TIMING
    hit_id = (uint32_t)(clock64() - timing_c0);
    rgba = timing_t0;
    dist = __uint_as_float(timer_ns_lo());
#endif
    if constexpr (STAGED) {
        stage[slot] = hit_id;
        stage[128u + slot] = rgba;
        stage[256u + slot] = __float_as_uint(dist);
    } else {
        f.hit_id[i] = hit_id;
        if (f.albedo) f.albedo[i] = rgba;  // nullptr: a gather peer shipping 8 B per pixel (kernels.cuh: FrameParams)
        f.distance[i] = dist;
        if (SHADE) f.shaded[i] = pixel;
    }
}

// ---- tile-sharded gather: flags between the GPUs that render one frame together (kernels.cuh: FrameParams) ----------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Root: the first CTA tells the peers that the framebuffer may be overwritten with frame `frame_seq`.
__device__ __forceinline__ void gather_prologue(const FrameParams& f) {
    if (f.go_flag != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) st_release_sys(f.go_flag, f.frame_seq);
}
// Peer: the last CTA to retire tells the root that this GPU's share of frame `frame_seq` is in the root's memory. The
// barrier orders every pixel store of the CTA before thread 0's system-scope fence; the CTA that finds the counter
// complete has observed all the others' increments (each made after such a fence) and fences again before the release.
__device__ __forceinline__ void gather_epilogue(const FrameParams& f) {
    if (f.done_flag == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (f.gather_tuning & GATHER_TUNE_CTA_FENCE_GPU)
            __threadfence();  // release at gpu scope towards the CTA that will publish; its system fence below is cumulative
        else
            __threadfence_system();
        if (atomicAdd(f.cta_counter, 1u) == gridDim.x * gridDim.y - 1u) {
            *f.cta_counter = 0u;  // launches of one view are ordered on one stream: the next one counts from zero
            __threadfence_system();
            st_release_sys(f.done_flag, f.frame_seq);
        }
    }
}

// Which 16x8-pixel block this CTA renders: its own position in the grid, or - heaviest-first order - the block the host's
// permutation assigns to its dispatch position (the hardware hands out CTAs in increasing linear index). Also starts the
// block's stopwatch when costs are recorded.
__device__ __forceinline__ void block_of_cta(const FrameParams& f, uint32_t& bx, uint32_t& by, long long* s_t0) {
    bx = blockIdx.x;
    by = blockIdx.y;
    if (f.cta_order != nullptr) {
        const uint32_t b = __ldg(f.cta_order + (blockIdx.y * gridDim.x + blockIdx.x));
        bx = b % gridDim.x;
        by = b / gridDim.x;
    }
    if (f.cta_cost != nullptr && threadIdx.x == 0) *s_t0 = clock64();
}
// ... and what the block cost, once its slowest warp is done
__device__ __forceinline__ void record_block_cost(const FrameParams& f, uint32_t bx, uint32_t by, const long long* s_t0) {
    if (f.cta_cost == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) f.cta_cost[by * gridDim.x + bx] = (uint32_t)min((long long)0xFFFFFFFFll, clock64() - *s_t0);
}

// Static schedule: one CTA per 16x8 pixel block of the frame. ORDERED = false: block = the CTA's place in the grid, nothing
// recorded (whole-frame views; exactly the round-1 kernels). ORDERED = true: heaviest-first order and / or cost recording
// (FrameParams::cta_order / cta_cost), used for the shards of a frame split over several GPUs.
template <bool LOD, int BS, bool ORDERED, bool SHARDED>
__device__ __forceinline__ void render_static_body(const DeviceTree& tree, const FrameParams& f) {
    int tx, ty;
    pixel_of_thread(tx, ty);
    gather_prologue(f);
    if constexpr (ORDERED) {
        __shared__ long long s_t0;
        uint32_t bx, by;
        block_of_cta(f, bx, by, &s_t0);
        shade_pixel<LOD, false, BS, false, SHARDED>(tree, f, bx * TILE_W + tx, by * TILE_H + ty);
        record_block_cost(f, bx, by, &s_t0);
    } else {
        shade_pixel<LOD, false, BS, false, SHARDED>(tree, f, blockIdx.x * TILE_W + tx, blockIdx.y * TILE_H + ty);
    }
    gather_epilogue(f);
}
// Instantiations: generic brick dimension (BS = -1) and the two the reference's examples use (8: cpu_render.rs:14, 32:
// dot_cube.rs:56, minecraft.rs:24, sponza.rs:24), where brick strides, masks and 1 / dim are immediates in the voxel loop;
// each without / with MIP maps (get_by_ray / get_by_ray_at_lod per pixel) and in raster / recorded order. Same code, same
// results; launch_render picks the instantiation.
#define SVX_RENDER_KERNEL(NAME, LOD, BS, ORDERED, SHARDED)                                                                  \
    __global__ void __launch_bounds__(BLOCK_THREADS, SVX_MIN_BLOCKS) NAME(const DeviceTree tree, const FrameParams f) {      \
        render_static_body<LOD, BS, ORDERED, SHARDED>(tree, f);                                                             \
    }
// whole frames
SVX_RENDER_KERNEL(render_kernel, false, -1, false, false)
SVX_RENDER_KERNEL(render_lod_kernel, true, -1, false, false)
SVX_RENDER_KERNEL(render_kernel_brick8, false, 3, false, false)
SVX_RENDER_KERNEL(render_kernel_brick32, false, 5, false, false)
SVX_RENDER_KERNEL(render_lod_kernel_brick8, true, 3, false, false)
SVX_RENDER_KERNEL(render_lod_kernel_brick32, true, 5, false, false)
// the rows of one rank of a frame split over several GPUs
SVX_RENDER_KERNEL(render_kernel_shard, false, -1, false, true)
SVX_RENDER_KERNEL(render_lod_kernel_shard, true, -1, false, true)
SVX_RENDER_KERNEL(render_kernel_shard_brick8, false, 3, false, true)
SVX_RENDER_KERNEL(render_kernel_shard_brick32, false, 5, false, true)
SVX_RENDER_KERNEL(render_lod_kernel_shard_brick8, true, 3, false, true)
SVX_RENDER_KERNEL(render_lod_kernel_shard_brick32, true, 5, false, true)
// ... and with the recorded block order / cost recording
SVX_RENDER_KERNEL(render_kernel_ordered, false, -1, true, true)
SVX_RENDER_KERNEL(render_lod_kernel_ordered, true, -1, true, true)
SVX_RENDER_KERNEL(render_kernel_ordered_brick8, false, 3, true, true)
SVX_RENDER_KERNEL(render_kernel_ordered_brick32, false, 5, true, true)
SVX_RENDER_KERNEL(render_lod_kernel_ordered_brick8, true, 3, true, true)
SVX_RENDER_KERNEL(render_lod_kernel_ordered_brick32, true, 5, true, true)
#undef SVX_RENDER_KERNEL
// Staged stores, for a gather peer: its pixels cross NVLink into rank 0's framebuffer, and what NVLink delivers depends on
// how they are written - 540 GB/s into one GPU for three planes as the 32-byte row segments of 8x4-pixel warp tiles, 755 GB/s
// as whole 128-byte lines (tools/nvlink_store_bench.cu, profiles/r02_nvlink_store_bench.json). So the CTA covers 32 x 4 pixels
// (four warp tiles side by side), every warp leaves its tile's values in shared memory, and after one barrier warp w writes
// row w of the block: 32 pixels, one 128-byte line per plane.
template <bool LOD, int BS>
__device__ __forceinline__ void render_staged_body(const DeviceTree& tree, const FrameParams& f) {
    __shared__ uint32_t stage[3 * 128];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    gather_prologue(f);
    {
        const uint32_t cx = (warp << 3) + (lane & 7u), cy = lane >> 3;  // this thread's pixel inside the block
        shade_pixel<LOD, false, BS, true>(tree, f, blockIdx.x * 32u + cx, blockIdx.y * 4u + cy, stage, cy * 32u + cx);
    }
    __syncthreads();
    const uint32_t x = blockIdx.x * 32u + lane, lr = blockIdx.y * 4u + warp;
    if (x < f.width && lr < f.rows_local) {
        const uint32_t row = image_row_of(f, lr);
        if (row < f.height) {
            const uint32_t i = (f.compact ? lr : row) * f.width + x, slot = warp * 32u + lane;
            f.hit_id[i] = stage[slot];
            if (f.albedo) f.albedo[i] = stage[128u + slot];
            f.distance[i] = __uint_as_float(stage[256u + slot]);
        }
    }
    gather_epilogue(f);
}
#define SVX_STAGED_KERNEL(NAME, LOD, BS)                                                                                    \
    __global__ void __launch_bounds__(128, SVX_MIN_BLOCKS) NAME(const DeviceTree tree, const FrameParams f) {               \
        render_staged_body<LOD, BS>(tree, f);                                                                               \
    }
SVX_STAGED_KERNEL(render_kernel_staged, false, -1)
SVX_STAGED_KERNEL(render_lod_kernel_staged, true, -1)
SVX_STAGED_KERNEL(render_kernel_staged_brick8, false, 3)
SVX_STAGED_KERNEL(render_kernel_staged_brick32, false, 5)
SVX_STAGED_KERNEL(render_lod_kernel_staged_brick8, true, 3)
SVX_STAGED_KERNEL(render_lod_kernel_staged_brick32, true, 5)
#undef SVX_STAGED_KERNEL

#ifndef SVX_BRICK_SPECIALISED
#define SVX_BRICK_SPECIALISED 1   // 0: always launch the generic kernels (A/B measurements, tools/probe_variants.sh)
#endif

// The same two with the shaded fourth plane (FrameParams::shaded), static schedule only
__global__ void __launch_bounds__(BLOCK_THREADS, SVX_MIN_BLOCKS) render_shaded_kernel(const DeviceTree tree, const FrameParams f) {
    int tx, ty;
    pixel_of_thread(tx, ty);
    gather_prologue(f);
    shade_pixel<false, true>(tree, f, blockIdx.x * TILE_W + tx, blockIdx.y * TILE_H + ty);
    gather_epilogue(f);
}
__global__ void __launch_bounds__(BLOCK_THREADS, SVX_MIN_BLOCKS) render_lod_shaded_kernel(const DeviceTree tree, const FrameParams f) {
    int tx, ty;
    pixel_of_thread(tx, ty);
    gather_prologue(f);
    shade_pixel<true, true>(tree, f, blockIdx.x * TILE_W + tx, blockIdx.y * TILE_H + ty);
    gather_epilogue(f);
}

// Persistent schedule: the grid is sized to the machine (SMs x resident CTAs) and every WARP pulls 8x4 pixel tiles from a
// global counter until the frame is done, so a long ray delays one warp, not the seven that share its CTA, and the SMs
// stay busy to the end of the frame. `counters[f.counter_slot]` is this launch's ticket counter; the other slot is
// zeroed for the next launch (launches of one view are ordered on one stream).
template <bool LOD, int BS = -1>
__device__ __forceinline__ void render_persistent_body(const DeviceTree& tree, const FrameParams& f, uint32_t* __restrict__ counters) {
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[f.counter_slot ^ 1u] = 0u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tiles_x = (f.width + 7u) >> 3, tiles_y = (f.rows_local + 3u) >> 2;
    const uint32_t blocks_x = (tiles_x + 3u) >> 2, blocks_y = (tiles_y + 1u) >> 1;
    // one ticket = SVX_TICKET_TILES consecutive tiles of a 32x8 pixel block (8 tiles): fewer atomics on the single
    // counter (64 800 tiles at 1080p would serialise in L2), still much finer than a CTA of the static schedule
    const uint32_t n_tickets = blocks_x * blocks_y * (8u / SVX_TICKET_TILES);
    for (;;) {
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(&counters[f.counter_slot], 1u);
        ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
        if (ticket >= n_tickets) return;
        const uint32_t first = ticket * SVX_TICKET_TILES;
        const uint32_t blk = first >> 3;
        const uint32_t bx = blk % blocks_x, by = blk / blocks_x;
#pragma unroll 1
        for (uint32_t k = 0; k < SVX_TICKET_TILES; ++k) {
            const uint32_t sub = (first & 7u) + k;
            const uint32_t ttx = (bx << 2) + (sub & 3u), tty = (by << 1) + (sub >> 2);
            if (ttx < tiles_x && tty < tiles_y) shade_pixel<LOD, false, BS>(tree, f, (ttx << 3) + (lane & 7u), (tty << 2) + (lane >> 3));
        }
    }
}
__global__ void __launch_bounds__(BLOCK_THREADS, SVX_MIN_BLOCKS) render_kernel_persistent(const DeviceTree tree, const FrameParams f,
                                                                                          uint32_t* __restrict__ counters) {
    gather_prologue(f);
    render_persistent_body<false>(tree, f, counters);
    gather_epilogue(f);
}
__global__ void __launch_bounds__(BLOCK_THREADS, SVX_MIN_BLOCKS) render_lod_kernel_persistent(const DeviceTree tree, const FrameParams f,
                                                                                              uint32_t* __restrict__ counters) {
    gather_prologue(f);
    render_persistent_body<true>(tree, f, counters);
    gather_epilogue(f);
}
// ... and for the brick dimensions of the reference's examples, like the static schedule above
#define SVX_PERSISTENT_KERNEL_FOR_BRICK(NAME, LOD, BS)                                                                      \
    __global__ void __launch_bounds__(BLOCK_THREADS, SVX_MIN_BLOCKS) NAME(const DeviceTree tree, const FrameParams f,        \
                                                                         uint32_t* __restrict__ counters) {                 \
        gather_prologue(f);                                                                                                 \
        render_persistent_body<LOD, BS>(tree, f, counters);                                                                 \
        gather_epilogue(f);                                                                                                 \
    }
SVX_PERSISTENT_KERNEL_FOR_BRICK(render_kernel_persistent_brick8, false, 3)
SVX_PERSISTENT_KERNEL_FOR_BRICK(render_kernel_persistent_brick32, false, 5)
SVX_PERSISTENT_KERNEL_FOR_BRICK(render_lod_kernel_persistent_brick8, true, 3)
SVX_PERSISTENT_KERNEL_FOR_BRICK(render_lod_kernel_persistent_brick32, true, 5)
#undef SVX_PERSISTENT_KERNEL_FOR_BRICK

// ---- Lane refill (EXPERIMENT, svx_view_set_schedule(view, 2); north-star subsystem 2: "ballot/shuffle-based ray compaction") ----
// Persistent warps whose lanes do not wait for the slowest ray of a tile: the traversal runs in rounds of f.refill_steps
// node-loop iterations (traverse_resumable, traverse_refill.cuh); after a round, lanes whose ray is done write their pixel,
// and once at least f.refill_min_idle lanes are idle they take the next pixels of the warp's pool (tiles pulled from the ticket
// counter, handed out in order with a ballot + popc rank) while the others keep their place in the
// tree. Framebuffer stores are per lane (scattered). Results are the same bits as every other schedule.

// What shade_pixel does before and after the traversal, split for the refill schedule.
// begin: 0 = the pixel is outside the frame, 1 = finished without a traversal (sky / certain miss: written here), 2 = a ray is on its way
template <bool LOD, int BS>
__device__ __forceinline__ int refill_begin_pixel(const DeviceTree& tree, const FrameParams& f, uint32_t x, uint32_t lr, uint32_t& pix,
                                                  RayConst& r, TraverseState& S) {
    if (x >= f.width || lr >= f.rows_local) return 0;
    const uint32_t row = image_row_of(f, lr);
    if (row >= f.height) return 0;
    const uint32_t y = f.height - 1u - row;
    pix = (f.compact ? lr : row) * f.width + x;
    bool may_hit = !(x < f.cull_x0 || x > f.cull_x1 || row < f.cull_row0 || row > f.cull_row1);
    if (may_hit) {
        float vx, vy, vz;
        glass_vector(f, x, y, vx, vy, vz);
        const float tree_size = (float)tree.tree_size;
        if (f.prefilter) {
            const float rl = rsqrtf((vx * vx) + (vy * vy) + (vz * vz));
            may_hit = !certain_root_miss(f.ox, f.oy, f.oz, vx * rl, vy * rl, vz * rl, tree_size);
        }
        if (may_hit) {
            const float len = sqrtf((vx * vx) + (vy * vy) + (vz * vz));
            r.ox = f.ox; r.oy = f.oy; r.oz = f.oz;
            r.dx = vx / len; r.dy = vy / len; r.dz = vz / len;
            float px, py, pz;
            uint32_t target_octant;
            may_hit = root_entry_and_setup(r, tree_size, px, py, pz, target_octant);
            if (may_hit) {
                traverse_begin<LOD, BS>(tree, S, px, py, pz, target_octant);
                return 2;
            }
        }
    }
    f.hit_id[pix] = NIL;
    if (f.albedo) f.albedo[pix] = 0u;
    f.distance[pix] = 0.0f;
    return 1;
}
__device__ __forceinline__ void refill_finish_pixel(const DeviceTree& tree, const FrameParams& f, uint32_t pix, const RayConst& r,
                                                    const TraceResult& res, bool hit) {
    uint32_t hit_id = NIL, rgba = 0u;
    float dist = 0.0f;
    if (hit) {
        hit_id = res.palette_value;
        const uint32_t ci = res.palette_value & 0xFFFFu;
        if (ci < 0xFFFFu && ci < tree.n_colors) rgba = __ldg(tree.palette + ci);
        const float wx = res.px - r.ox, wy = res.py - r.oy, wz = res.pz - r.oz;
        dist = sqrtf((wx * wx) + (wy * wy) + (wz * wz));  // V3c::length, vector.rs:75-77
    }
    f.hit_id[pix] = hit_id;
    if (f.albedo) f.albedo[pix] = rgba;
    f.distance[pix] = dist;
}

template <bool LOD, int BS>
__device__ __forceinline__ void render_refill_body(const DeviceTree& tree, const FrameParams& f, uint32_t* __restrict__ counters) {
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[f.counter_slot ^ 1u] = 0u;
    const uint32_t lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
    const uint32_t tiles_x = (f.width + 7u) >> 3, tiles_y = (f.rows_local + 3u) >> 2;
    const uint32_t blocks_x = (tiles_x + 3u) >> 2, blocks_y = (tiles_y + 1u) >> 1, n_blocks = blocks_x * blocks_y;
    // the warp's pool: f.refill_unit_tiles (1, 2, 4 or 8) consecutive 8x4 tiles of a 32x8-pixel block per ticket, `first_tile`
    // the first of them, pixels q = next_q .. pool_size-1 still to hand out
    const uint32_t unit = f.refill_unit_tiles, pool_size = unit << 5, n_tickets = n_blocks * (8u / unit);
    uint32_t first_tile = 0u, next_q = pool_size;
    bool pool_empty = false, has_ray = false;
    uint32_t pix = 0u;
    RayConst r;
    TraverseState S;
    for (;;) {
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !has_ray);
        if (!pool_empty && (idle == 0xFFFFFFFFu || (uint32_t)__popc(idle) >= f.refill_min_idle)) {
            uint32_t want = idle;
            while (want != 0u) {
                if (next_q >= pool_size) {
                    uint32_t ticket = 0u;
                    if (lane == 0) ticket = atomicAdd(&counters[f.counter_slot], 1u);
                    ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
                    if (ticket >= n_tickets) {
                        pool_empty = true;
                        break;
                    }
                    first_tile = ticket * unit;
                    next_q = 0u;
                }
                const uint32_t take = min((uint32_t)__popc(want), pool_size - next_q), rank = (uint32_t)__popc(want & below);
                const bool served = ((want >> lane) & 1u) != 0u && rank < take;
                if (served) {
                    const uint32_t q = next_q + rank, tile = first_tile + (q >> 5), in_tile = q & 31u, blk = tile >> 3, sub = tile & 7u;
                    const uint32_t ttx = ((blk % blocks_x) << 2) + (sub & 3u), tty = ((blk / blocks_x) << 1) + (sub >> 2);
                    has_ray = refill_begin_pixel<LOD, BS>(tree, f, (ttx << 3) + (in_tile & 7u), (tty << 2) + (in_tile >> 3), pix, r, S) == 2;
                }
                want &= ~__ballot_sync(0xFFFFFFFFu, served);
                next_q += take;
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, has_ray) == 0u) {
            if (pool_empty) return;
            continue;
        }
        if (has_ray) {
            TraceResult res;
            const int walk = traverse_resumable<LOD, BS>(tree, r, S, res, f.viewing_distance, f.refill_steps);
            if (walk != WALK_SUSPENDED) {
                refill_finish_pixel(tree, f, pix, r, res, walk == WALK_HIT);
                has_ray = false;
            }
        }
    }
}
#define SVX_REFILL_KERNEL(NAME, LOD, BS)                                                                                    \
    __global__ void __launch_bounds__(BLOCK_THREADS, SVX_MIN_BLOCKS) NAME(const DeviceTree tree, const FrameParams f,        \
                                                                         uint32_t* __restrict__ counters) {                 \
        gather_prologue(f);                                                                                                 \
        render_refill_body<LOD, BS>(tree, f, counters);                                                                     \
        gather_epilogue(f);                                                                                                 \
    }
SVX_REFILL_KERNEL(render_kernel_refill, false, -1)
SVX_REFILL_KERNEL(render_kernel_refill_brick8, false, 3)
SVX_REFILL_KERNEL(render_kernel_refill_brick32, false, 5)
SVX_REFILL_KERNEL(render_lod_kernel_refill, true, -1)
SVX_REFILL_KERNEL(render_lod_kernel_refill_brick8, true, 3)
SVX_REFILL_KERNEL(render_lod_kernel_refill_brick32, true, 5)
#undef SVX_REFILL_KERNEL

template <bool LOD>
__device__ __forceinline__ void rays_body(const DeviceTree& tree, const float* __restrict__ rays, uint64_t n,
                                          float viewing_distance, RayHitRecord* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayConst r;
    r.ox = rays[6 * i + 0]; r.oy = rays[6 * i + 1]; r.oz = rays[6 * i + 2];
    r.dx = rays[6 * i + 3]; r.dy = rays[6 * i + 4]; r.dz = rays[6 * i + 5];
    ray_setup(r);
    TraceResult res;
    const bool hit = trace_ray<LOD>(tree, r, res, viewing_distance);
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
__global__ void __launch_bounds__(BLOCK_THREADS, 6) rays_kernel(const DeviceTree tree, const float* __restrict__ rays,
                                                             uint64_t n, RayHitRecord* __restrict__ out) {
    rays_body<false>(tree, rays, n, 0.0f, out);
}
__global__ void __launch_bounds__(BLOCK_THREADS, 6) rays_lod_kernel(const DeviceTree tree, const float* __restrict__ rays,
                                                                 uint64_t n, float viewing_distance,
                                                                 RayHitRecord* __restrict__ out) {
    rays_body<true>(tree, rays, n, viewing_distance, out);
}

// pix_points_to_empty (reference src/octree/node.rs:405-427) negated: does this palette value show or carry anything?
// `shows` / `carries` are bit tables in shared memory over the WHOLE u16 index range (2048 words each): bit c of `shows`
// = colour c exists and has albedo.a != 0, bit d of `carries` = user data d exists and is not zero. Indices beyond the
// palettes, and 0xFFFF = "none", read a 0 bit, so the test needs no compare and no branch.
constexpr uint32_t TABLE_WORDS = 65536u / 32u;
__device__ __forceinline__ bool voxel_occupied(const uint32_t* shows, const uint32_t* carries, uint32_t v) {
    const uint32_t ci = v & 0xFFFFu, di = v >> 16;
    return (((shows[ci >> 5] >> (ci & 31u)) | (carries[di >> 5] >> (di & 31u))) & 1u) != 0u;
}

// Both kernels start by staging the two bit tables in shared memory (`tables`, host-built = color_words words, then
// data_words words; the rest of each 2048-word table is zero).
__device__ __forceinline__ void stage_tables(uint32_t* smem, const uint32_t* __restrict__ tables, uint32_t color_words,
                                             uint32_t data_words) {
    for (uint32_t i = threadIdx.x; i < TABLE_WORDS; i += blockDim.x) {
        smem[i] = i < color_words ? __ldg(tables + i) : 0u;
        smem[TABLE_WORDS + i] = i < data_words ? __ldg(tables + color_words + i) : 0u;
    }
    __syncthreads();
}

// Occupancy bit-bricks of the listed bricks. HBM-bound streaming kernel: 4 B read per voxel, 1 bit written.
// Bricks of >= 32 voxels (brick_dim >= 4): a thread loads 4 voxels with one 16-byte load, 8 neighbouring lanes make one
// 32-voxel word (butterfly OR of their nibbles), so a warp turns 512 contiguous bytes into 4 words per step; all
// index arithmetic is shifts (volume and words per brick are powers of two).
__global__ void __launch_bounds__(256) occupancy_bits_kernel(const DeviceTree tree, const uint32_t* __restrict__ tables,
                                                             uint32_t color_words, uint32_t data_words,
                                                             const uint32_t* __restrict__ handles, uint32_t n,
                                                             uint32_t* __restrict__ bits_out) {
    __shared__ uint32_t smem[2 * TABLE_WORDS];
    stage_tables(smem, tables, color_words, data_words);
    const uint32_t* shows = smem;
    const uint32_t* carries = smem + TABLE_WORDS;
    const uint32_t lane = threadIdx.x & 31u, sub = lane & 7u;
    const uint32_t vol_shift = 3u * tree.brick_shift;  // log2(voxels per brick) >= 5 here
    const uint32_t word_shift = vol_shift - 5u;        // log2(words per brick)
    // 32-bit word counters: the host guarantees n * words_per_brick < 2^32 (capi.cu: upload)
    const uint32_t total_words = n << word_shift;
    const uint32_t groups = (gridDim.x * blockDim.x) >> 3;  // 8-lane groups in the grid, one word each per step
    constexpr int UNROLL = 4;  // independent 16-byte loads in flight per thread (64 B): covers HBM latency at this occupancy
    const uint32_t trips = (uint32_t)(((uint64_t)total_words + (uint64_t)groups * UNROLL - 1) / ((uint64_t)groups * UNROLL));
    const uint32_t word_mask = (1u << word_shift) - 1u;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    for (uint32_t t = 0; t < trips; ++t) {  // same trip count for every lane: the shuffles stay converged
        uint4 v[UNROLL];
        uint32_t out_index[UNROLL];  // word index into bits_out; fits 32 bits by the same guarantee
        bool valid[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t w64 = ((uint64_t)t * UNROLL + (uint64_t)u) * groups + g;
            valid[u] = w64 < total_words;
            const uint32_t w = (uint32_t)w64;
            v[u] = make_uint4(NIL, NIL, NIL, NIL);
            out_index[u] = 0u;
            if (valid[u]) {
                const uint32_t handle = __ldg(handles + (w >> word_shift));
                const uint32_t word = w & word_mask;
                out_index[u] = (handle << word_shift) + word;
                v[u] = __ldg(reinterpret_cast<const uint4*>(tree.voxels + ((size_t)handle << vol_shift)) + ((word << 3) + sub));
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            // voxel data is coherent (air, single-material runs): four equal palette values need one table test
            uint32_t nib;
            if (v[u].x == v[u].y && v[u].z == v[u].w && v[u].x == v[u].z)
                nib = voxel_occupied(shows, carries, v[u].x) ? 0xFu : 0u;
            else
                nib = (voxel_occupied(shows, carries, v[u].x) ? 1u : 0u) | (voxel_occupied(shows, carries, v[u].y) ? 2u : 0u) |
                      (voxel_occupied(shows, carries, v[u].z) ? 4u : 0u) | (voxel_occupied(shows, carries, v[u].w) ? 8u : 0u);
            uint32_t part = nib << (4u * sub);  // NIL voxels (lanes past the end) are never occupied
            part |= __shfl_xor_sync(0xFFFFFFFFu, part, 1);
            part |= __shfl_xor_sync(0xFFFFFFFFu, part, 2);
            part |= __shfl_xor_sync(0xFFFFFFFFu, part, 4);
            if (sub == 0u && valid[u]) bits_out[out_index[u]] = part;
        }
    }
}

// The same for bricks of fewer than 32 voxels (brick_dim 1 or 2: one partly used word per brick), one warp per brick.
__global__ void __launch_bounds__(256) occupancy_bits_small_kernel(const DeviceTree tree, const uint32_t* __restrict__ tables,
                                                                   uint32_t color_words, uint32_t data_words,
                                                                   const uint32_t* __restrict__ handles, uint32_t n,
                                                                   uint32_t* __restrict__ bits_out) {
    __shared__ uint32_t smem[2 * TABLE_WORDS];
    stage_tables(smem, tables, color_words, data_words);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t vol = 1u << (3u * tree.brick_shift);
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n; b += warps) {  // warp-uniform bounds
        const uint32_t handle = __ldg(handles + b);
        const bool occupied = lane < vol && voxel_occupied(smem, smem + TABLE_WORDS, __ldg(tree.voxels + (size_t)handle * vol + lane));
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, occupied);
        if (lane == 0) bits_out[handle] = mask;
    }
}

// Evaluates the closed forms that replace the reference's tables, for every table index.
__global__ void lut_selftest_kernel(uint64_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 512u) {
        // RAY_TO_NODE_OCCUPANCY_BITMASK_LUT[cell][dir]: recover the mask bit by bit from the predicate
        const uint32_t cell = i >> 3, dir = i & 7u;
        const uint32_t cx = cell & 3u, cy = (cell >> 2) & 3u, cz = cell >> 4;
        uint64_t m = 0;
        for (uint32_t b = 0; b < 64u; ++b) {
            const uint64_t one = 1ull << b;
            if (ray_may_hit((uint32_t)one, (uint32_t)(one >> 32), cx, cy, cz, dir)) m |= one;
        }
        out[i] = m;
    } else if (i < 520u) {
        const uint32_t o = i - 512u;
        uint64_t m = 0;
        for (uint32_t b = 0; b < 64u; ++b) {
            const uint64_t one = 1ull << b;
            if (octant_occupied((uint32_t)one, (uint32_t)(one >> 32), o)) m |= one;
        }
        out[i] = m;
    } else if (i < 520u + 216u) {
        const uint32_t k = i - 520u;
        const uint32_t o = k & 7u, s = k >> 3;  // s = (x+1)*9 + (y+1)*3 + (z+1)
        const int sx = (int)(s / 9u) - 1, sy = (int)((s / 3u) % 3u) - 1, sz = (int)(s % 3u) - 1;
        // per axis: stepped = the step component is not 0, its sign as a posbit (irrelevant for an axis that did not step)
        const uint32_t posbits = (sx > 0 ? 1u : 0u) | (sz > 0 ? 2u : 0u) | (sy > 0 ? 4u : 0u);
        out[i] = step_octant(o, sx != 0, sy != 0, sz != 0, posbits);
    }
}

cudaError_t launch_render(const DeviceTree& tree, const FrameParams& frame, const LaunchConfig& cfg, cudaStream_t stream) {
    if (frame.rows_local == 0 || frame.width == 0) return cudaSuccess;
    if (frame.shaded) {  // the shaded plane exists in the static schedule only
        dim3 grid((frame.width + TILE_W - 1) / TILE_W, (frame.rows_local + TILE_H - 1) / TILE_H);
        if (tree.mips_enabled)
            render_lod_shaded_kernel<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame);
        else
            render_shaded_kernel<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame);
        return cudaGetLastError();
    }
    // the specialised instantiations hard-code everything DeviceTree derives from brick_shift; anything else is generic
    const bool consistent = tree.brick_dim == (1u << tree.brick_shift) && tree.brick_dim_sq == tree.brick_dim * tree.brick_dim;
    const uint32_t shift = (SVX_BRICK_SPECIALISED && consistent) ? tree.brick_shift : 0xFFFFFFFFu;
    if (cfg.persistent && cfg.refill && cfg.tile_counters) {
        const unsigned grid = (unsigned)(cfg.sm_count * SVX_MIN_BLOCKS);
        if (tree.mips_enabled) {
            if (shift == 3u) render_lod_kernel_refill_brick8<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
            else if (shift == 5u) render_lod_kernel_refill_brick32<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
            else render_lod_kernel_refill<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
        } else {
            if (shift == 3u) render_kernel_refill_brick8<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
            else if (shift == 5u) render_kernel_refill_brick32<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
            else render_kernel_refill<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
        }
        return cudaGetLastError();
    }
    if (cfg.persistent && cfg.tile_counters) {
        // blocks_x * blocks_y * 8 tickets cover the frame in 32x8 blocks; ragged edges are skipped inside the kernel
        const unsigned grid = (unsigned)(cfg.sm_count * SVX_MIN_BLOCKS);
        if (tree.mips_enabled) {
            if (shift == 3u) render_lod_kernel_persistent_brick8<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
            else if (shift == 5u) render_lod_kernel_persistent_brick32<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
            else render_lod_kernel_persistent<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
        } else {
            if (shift == 3u) render_kernel_persistent_brick8<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
            else if (shift == 5u) render_kernel_persistent_brick32<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
            else render_kernel_persistent<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame, cfg.tile_counters);
        }
        return cudaGetLastError();
    }
    if (cfg.staged_stores && !frame.shaded) {
        dim3 sgrid((frame.width + 31u) / 32u, (frame.rows_local + 3u) / 4u);
#define SVX_LAUNCH(K) K<<<sgrid, 128, 0, stream>>>(tree, frame)
        if (tree.mips_enabled) {
            if (shift == 3u) SVX_LAUNCH(render_lod_kernel_staged_brick8);
            else if (shift == 5u) SVX_LAUNCH(render_lod_kernel_staged_brick32);
            else SVX_LAUNCH(render_lod_kernel_staged);
        } else {
            if (shift == 3u) SVX_LAUNCH(render_kernel_staged_brick8);
            else if (shift == 5u) SVX_LAUNCH(render_kernel_staged_brick32);
            else SVX_LAUNCH(render_kernel_staged);
        }
#undef SVX_LAUNCH
        return cudaGetLastError();
    }
    dim3 grid((frame.width + TILE_W - 1) / TILE_W, (frame.rows_local + TILE_H - 1) / TILE_H);
    const bool ordered = frame.cta_order != nullptr || frame.cta_cost != nullptr, sharded = frame.world != 1u;
#define SVX_LAUNCH(K) K<<<grid, BLOCK_THREADS, 0, stream>>>(tree, frame)
#define SVX_PICK(STEM, SUFFIX)                                                \
    do {                                                                      \
        if (ordered) SVX_LAUNCH(STEM##_ordered##SUFFIX);                      \
        else if (sharded) SVX_LAUNCH(STEM##_shard##SUFFIX);                   \
        else SVX_LAUNCH(STEM##SUFFIX);                                        \
    } while (0)
    if (tree.mips_enabled) {
        if (shift == 3u) SVX_PICK(render_lod_kernel, _brick8);
        else if (shift == 5u) SVX_PICK(render_lod_kernel, _brick32);
        else SVX_PICK(render_lod_kernel, );
    } else {
        if (shift == 3u) SVX_PICK(render_kernel, _brick8);
        else if (shift == 5u) SVX_PICK(render_kernel, _brick32);
        else SVX_PICK(render_kernel, );
    }
#undef SVX_PICK
#undef SVX_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_rays(const DeviceTree& tree, const float* rays, uint64_t n, float viewing_distance, RayHitRecord* out,
                        const LaunchConfig&, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + BLOCK_THREADS - 1) / BLOCK_THREADS);
    if (tree.mips_enabled)
        rays_lod_kernel<<<blocks, BLOCK_THREADS, 0, stream>>>(tree, rays, n, viewing_distance, out);
    else
        rays_kernel<<<blocks, BLOCK_THREADS, 0, stream>>>(tree, rays, n, out);
    return cudaGetLastError();
}

cudaError_t launch_occupancy_bits(const DeviceTree& tree, const uint32_t* tables, uint32_t color_words, uint32_t data_words,
                                  const uint32_t* handles, uint32_t n, uint32_t* bits_out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (color_words > 2048u || data_words > 2048u) return cudaErrorInvalidValue;  // u16 palette indices (types.rs:188-191)
    const unsigned machine = 148u * 8u * 2u;  // two waves of the 8 CTAs an SM holds (32 registers x 256 threads), then grid-stride
    if (tree.brick_shift < 2u) {
        const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)n + 7) / 8, machine);
        occupancy_bits_small_kernel<<<grid, 256, 0, stream>>>(tree, tables, color_words, data_words, handles, n, bits_out);
    } else {
        const uint64_t total_words = (uint64_t)n * tree.bit_words;
        const unsigned grid = (unsigned)std::min<uint64_t>((total_words + 31) / 32, machine);  // 32 words per CTA and step
        occupancy_bits_kernel<<<grid, 256, 0, stream>>>(tree, tables, color_words, data_words, handles, n, bits_out);
    }
    return cudaGetLastError();
}

// One thread waits until *flag has reached `want` (sequence numbers: compared modulo 2^32). The flag may live in another
// GPU's memory (a peer polling the root's `go` over NVLink): the poll backs off between reads. A flag that does not
// arrive within timeout_ns is reported through the host-mapped error word instead of hanging the device.
__global__ void wait_flag_kernel(const uint32_t* flag, uint32_t want, uint64_t timeout_ns, uint32_t* error, uint32_t error_code) {
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - want) < 0) {
        if (globaltimer_ns() - t0 > timeout_ns) {
            *error = error_code;
            __threadfence_system();
            return;
        }
        __nanosleep(100);
    }
}

// Root side of the gather: every CTA waits until all peers have published `frame_seq` in their done word (their pixels are
// then in this GPU's memory), and - for the 8-byte wire format - resolves the albedo of the peers' rows from the hit ids
// exactly as the viewport kernel does for its own pixels (palette[hit_id & 0xFFFF], 0 without a colour).
__global__ void __launch_bounds__(256) gather_complete_kernel(const GatherComplete g) {
    __shared__ uint32_t arrived_mask, failed;
    if (threadIdx.x == 0) {
        arrived_mask = 1u;  // rank 0 is this GPU
        failed = 0u;
    }
    __syncthreads();
    const uint64_t t0 = globaltimer_ns();
    // thread 0 waits for peer r (once per CTA); everybody learns the outcome through shared memory
    auto wait_for = [&](uint32_t r) -> bool {
        const bool need = !((arrived_mask >> r) & 1u);
        __syncthreads();  // everybody has read the mask before thread 0 may change it
        if (need) {
            if (threadIdx.x == 0) {
                const uint32_t* flag = g.done_flags + (size_t)r * g.done_stride;
                bool ok = true;
                while ((int32_t)(ld_acquire_sys(flag) - g.frame_seq) < 0) {
                    if (globaltimer_ns() - t0 > g.timeout_ns) {
                        ok = false;
                        break;
                    }
                    __nanosleep(40);
                }
                if (ok) {
                    arrived_mask |= 1u << r;
                } else {
                    failed = 1u;
                    if (blockIdx.x == 0 || g.fill_albedo) *g.error = 1u + r;
                    __threadfence_system();
                }
            }
            __syncthreads();
        }
        return failed == 0u;
    };
    if (!g.fill_albedo) {  // one CTA: the frame is complete when every peer has published this frame
        for (uint32_t r = 1; r < g.world; ++r)
            if (!wait_for(r)) return;
        return;
    }
    // 8-byte wire format: a row can be finished as soon as ITS owner has delivered, so the albedo of early peers' rows is
    // resolved while late peers are still rendering. Every peer row belongs to exactly one CTA, so when the grid has
    // retired every peer has been waited for.
    const bool vec = (g.width & 3u) == 0u;
    for (uint32_t row = blockIdx.x; row < g.height; row += gridDim.x) {
        const uint32_t band = row >> g.band_shift;
        const uint32_t owner = g.band_rotate ? (band % g.world + band / g.world) % g.world : band % g.world;
        if (owner == 0u) continue;  // the root's own rows already carry their albedo
        if (!wait_for(owner)) return;
        const size_t base = (size_t)row * g.width;
        if (vec) {
            const uint4* src = reinterpret_cast<const uint4*>(g.hit_id + base);
            uint4* dst = reinterpret_cast<uint4*>(g.albedo + base);
            for (uint32_t x = threadIdx.x; x < (g.width >> 2); x += blockDim.x) {
                const uint4 h = __ldcg(src + x);
                uint4 a;
                const uint32_t c0 = h.x & 0xFFFFu, c1 = h.y & 0xFFFFu, c2 = h.z & 0xFFFFu, c3 = h.w & 0xFFFFu;
                a.x = (c0 < 0xFFFFu && c0 < g.n_colors) ? __ldg(g.palette + c0) : 0u;
                a.y = (c1 < 0xFFFFu && c1 < g.n_colors) ? __ldg(g.palette + c1) : 0u;
                a.z = (c2 < 0xFFFFu && c2 < g.n_colors) ? __ldg(g.palette + c2) : 0u;
                a.w = (c3 < 0xFFFFu && c3 < g.n_colors) ? __ldg(g.palette + c3) : 0u;
                dst[x] = a;
            }
        } else {
            for (uint32_t x = threadIdx.x; x < g.width; x += blockDim.x) {
                const uint32_t c = __ldcg(g.hit_id + base + x) & 0xFFFFu;
                g.albedo[base + x] = (c < 0xFFFFu && c < g.n_colors) ? __ldg(g.palette + c) : 0u;
            }
        }
    }
}

// Heaviest-first permutation of the blocks of a launch (FrameParams::cta_order) from the cycles each took in the frame that
// just ended: a counting sort over 64 quarter-octave cost classes by ONE CTA (a share of a 4K frame has 8 100 - 32 400
// blocks, a whole one 64 800). Within a class the blocks keep (roughly) their raster order, which keeps neighbours together.
constexpr uint32_t ORDER_CLASSES = 64;
__device__ __forceinline__ uint32_t block_cost_class(uint32_t cycles) {
    cycles = max(cycles, 4u);
    const uint32_t lz = 31u - (uint32_t)__clz((int)cycles);
    const uint32_t level = 4u * lz + ((cycles >> (lz - 2u)) & 3u);  // floor(4 log2 cycles)
    return level >= 96u ? 0u : min(96u - level, ORDER_CLASSES - 1u);  // class 0: >= 2^24 cycles; 16x8 blocks sit around 2^14 - 2^20
}
__global__ void __launch_bounds__(1024) order_ctas_kernel(const uint32_t* __restrict__ cost, uint32_t* __restrict__ order, uint32_t n, uint32_t head) {
    __shared__ uint32_t cursor[ORDER_CLASSES];
    __shared__ uint32_t s_cut, s_head_total, s_warp_sums[32];
    if (threadIdx.x < ORDER_CLASSES) cursor[threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&cursor[block_cost_class(__ldcg(cost + i))], 1u);
    __syncthreads();
    if (threadIdx.x < 32u) {  // exclusive prefix over the classes, heaviest first: two classes per lane, then a warp scan
        const uint32_t a = cursor[2u * threadIdx.x], b = cursor[2u * threadIdx.x + 1u];
        uint32_t incl = a + b;
#pragma unroll
        for (uint32_t d = 1; d < 32u; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (threadIdx.x >= d) incl += up;
        }
        // the HEAD of the order: whole classes, heaviest first, until at least `head` blocks are in it; everything lighter
        // keeps its raster order behind them (a fully sorted frame ends in a burst of short blocks - and of their stores, which
        // for a gather peer all cross NVLink at once)
        const bool a_in = incl - a - b < head, b_in = incl - b < head;
        const uint32_t votes = __ballot_sync(0xFFFFFFFFu, b_in), votes_a = __ballot_sync(0xFFFFFFFFu, a_in);
        if (threadIdx.x == 0) {
            const uint32_t full = (uint32_t)__popc(votes);  // lanes whose both classes are in the head
            s_cut = 2u * full + ((full < 32u && ((votes_a >> full) & 1u)) ? 1u : 0u);  // classes [0, s_cut) form the head
        }
        cursor[2u * threadIdx.x] = incl - a - b;
        cursor[2u * threadIdx.x + 1u] = incl - b;
        __syncwarp();
        if (threadIdx.x == 0) {
            const uint32_t cut = s_cut;
            s_head_total = cut >= ORDER_CLASSES ? n : cursor[cut];
        }
    }
    __syncthreads();
    const uint32_t cut = s_cut;
    // head: scattered by class; tail: a stable compaction, 1024 blocks per pass (ballot per warp, the 32 warp counts scanned by
    // every warp with shuffles, the running offset kept in a register)
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t tail_at = s_head_total;
    uint32_t ahead = threadIdx.x < n ? __ldcg(cost + threadIdx.x) : 0u;  // the next pass' cost is in flight while this pass synchronises
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t cls = i < n ? block_cost_class(ahead) : 0u;
        if (i + blockDim.x < n) ahead = __ldcg(cost + i + blockDim.x);
        const bool tail = i < n && cls >= cut;
        if (i < n && !tail) order[atomicAdd(&cursor[cls], 1u)] = i;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, tail);
        if (lane == 0) s_warp_sums[warp] = (uint32_t)__popc(mask);
        __syncthreads();
        const uint32_t mine = s_warp_sums[lane];
        uint32_t incl = mine;
#pragma unroll
        for (uint32_t d = 1; d < 32u; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += up;
        }
        const uint32_t before = __shfl_sync(0xFFFFFFFFu, incl - mine, warp), total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (tail) order[tail_at + before + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = i;
        tail_at += total;
        __syncthreads();
    }
}

cudaError_t launch_order_ctas(const uint32_t* cost, uint32_t* order, uint32_t n, uint32_t head, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    order_ctas_kernel<<<1, 1024, 0, stream>>>(cost, order, n, head);
    return cudaGetLastError();
}

__global__ void signal_flag_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    st_release_sys(flag, value);
}

cudaError_t launch_signal_flag(uint32_t* flag, uint32_t value, cudaStream_t stream) {
    signal_flag_kernel<<<1, 1, 0, stream>>>(flag, value);
    return cudaGetLastError();
}

cudaError_t launch_wait_flag(const uint32_t* flag, uint32_t want, uint64_t timeout_ns, uint32_t* error, uint32_t error_code,
                             cudaStream_t stream) {
    wait_flag_kernel<<<1, 1, 0, stream>>>(flag, want, timeout_ns, error, error_code);
    return cudaGetLastError();
}

cudaError_t launch_gather_complete(const GatherComplete& g, int sm_count, cudaStream_t stream) {
    // waiting only: one CTA. Filling: two CTAs per SM stream the peers' rows (HBM-bound, 4 B read + 4 B written per pixel)
    const unsigned grid = g.fill_albedo ? (unsigned)std::min<uint32_t>(g.height, (uint32_t)sm_count * 2u) : 1u;
    gather_complete_kernel<<<grid, 256, 0, stream>>>(g);
    return cudaGetLastError();
}

cudaError_t launch_lut_selftest(uint64_t* out, cudaStream_t stream) {
    lut_selftest_kernel<<<(736 + 127) / 128, 128, 0, stream>>>(out);
    return cudaGetLastError();
}

}  // namespace svx
