// Kernel entry points of the primary-ray path (launch wrappers are in kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gpu_tree.hpp"

namespace svx {

// Per-frame camera constants of the caller loop in reference examples/cpu_render.rs:78-103, computed once on the
// host in the reference's f32 operation order (capi.cu: make_frame_constants).
struct FrameParams {
    float ox, oy, oz;      // viewport_ray.origin
    float blx, bly, blz;   // viewport_bottom_left
    float rx, ry, rz;      // viewport_right_direction
    float ux, uy, uz;      // viewport_up_direction = (0, 1, 0)
    float pixel_width, pixel_height;
    uint32_t width, height;
    // sharding: this launch renders image rows r with (r / band_rows) % world == rank
    uint32_t rank, world, band_shift;  // bands of 2^band_shift rows
    // 1: rotated interleave (gather members) - band g belongs to rank (g % world + g / world) % world, so the rank order
    // shifts by one in every cycle of `world` bands and a cost pattern whose period on the screen matches the cycle (regular
    // architecture, 64 rows at 8 GPUs) does not land on the same rank every time: at 8 GPUs one rank's share of sponza was
    // 7 % heavier than the others' with the plain interleave (profiles/r02_gather_probe_n8_staged.json)
    uint32_t band_rotate;
    uint32_t world_magic;  // ceil(2^32 / world): band % world = band - umulhi(band, world_magic) * world, exact for band < 2^16
    uint32_t rows_local;   // number of image rows this shard owns
    // pixels outside [cull_x0, cull_x1] x [cull_row0, cull_row1] (image coordinates, inclusive) certainly miss the root
    // cube: a conservative screen-space bound of the cube's projection computed on the host (capi.cu)
    uint32_t cull_x0, cull_x1, cull_row0, cull_row1;
    uint32_t prefilter;    // 1: run the approximate root-miss test per ray (no cull rectangle available for this pose)
    uint32_t counter_slot; // persistent schedule: which of the two ticket counters this launch consumes
    uint32_t compact;      // 1: store shard-local row lr at output row lr (band-major compact buffer for gathers)
    float viewing_distance;  // get_by_ray_at_lod's parameter (raytracing_on_cpu.rs:325); only read when tree.mips_enabled
    uint32_t* hit_id;      // [h*w]
    uint32_t* albedo;      // [h*w]
    float* distance;       // [h*w]
    // optional fourth plane: the pixel the reference's caller loops write (examples/cpu_render.rs:119-136) - albedo
    // scaled by a diffuse term from the impact normal and this light normal, grey on a miss. nullptr = not produced.
    uint32_t* shaded;      // [h*w] RGBA8, r in the low byte
    float lx, ly, lz;      // diffuse_light_normal (cpu_render.rs:97)
    // Tile-sharded gather (multi_gpu.cu): one frame rendered by `world` GPUs into the ROOT GPU's framebuffer, the peers'
    // stores crossing NVLink from inside this kernel. All three are nullptr outside a gather.
    //   go_flag    root only: the first CTA publishes `frame_seq` here (release, system scope) - the peers' licence to
    //              overwrite the root's framebuffer with frame `frame_seq` (everything the consumer queued before this
    //              launch has retired by then)
    //   done_flag  peers only: the root's done[rank] word, written with `frame_seq` by the last CTA of this launch to
    //              retire, after a system-scope fence: all of this GPU's pixels are in the root's memory
    //   cta_counter peers only: local counter of retired CTAs (reset by the CTA that finds it complete)
    // A peer launched with albedo == nullptr ships 8 B per pixel (hit id + distance); the root fills the albedo of the
    // peers' rows from its own palette (gather_complete_kernel).
    // Heaviest-first order of the static schedule. The end of a launch is a tail of a few long rays during which most of the
    // machine idles - 23 us of sponza 4K's 840 us, and the same ~30 us of the 130 us ONE of eight GPUs spends on its share of
    // that frame (profiles/r02_pixel_timing_sponza.json): it is what bounds strong scaling. So the 16x8-pixel blocks are
    // dispatched heaviest first, by the cycles they took in the PREVIOUS frame of the view (screen-space cost changes slowly
    // between frames): every CTA records its block's cycles in cta_cost, a one-CTA counting sort queued behind the frame
    // (order_ctas_kernel) turns them into the next frame's permutation, and the i-th CTA the hardware dispatches renders
    // block cta_order[i]. The traversal code itself is untouched: one broadcast load per CTA. nullptr = raster order / no
    // recording (first frame, whole-frame views by default).
    const uint32_t* cta_order;  // [ctas] permutation of the block ids
    uint32_t* cta_cost;         // [ctas] cycles per block, written by this launch
    uint32_t* go_flag;
    uint32_t* done_flag;
    uint32_t* cta_counter;
    uint32_t frame_seq;
    uint32_t gather_tuning;  // GATHER_TUNE_* bits (multi_gpu.cu), 0 = the defaults
    // lane-refill schedule only: node-loop iterations per round, and how many lanes must be idle before the warp hands out pixels
    uint32_t refill_steps, refill_min_idle, refill_unit_tiles;  // ... and 8x4 tiles per ticket (1, 2, 4, 8)
};

// Tuning bits of the gather (SVX_GATHER_TUNING, read when a view opens / joins a gather). Default 0 = what measured best at
// 2, 4 and 8 GPUs (profiles/r02_gather_probe_n*.json): peers run the static schedule and publish `done` from a one-thread
// kernel queued behind the viewport kernel - the kernel boundary orders the stores, no fence inside the viewport kernel
// (a membar.sys per 16x8-pixel block cost 18 % of the peer's kernel even with local stores).
constexpr uint32_t GATHER_TUNE_CTA_FENCE_GPU = 1u;    // in-kernel variants: retiring CTAs fence at gpu scope, only the publishing one at system scope
constexpr uint32_t GATHER_TUNE_INKERNEL_STATIC = 2u;  // peers: static schedule, the viewport kernel's last CTA publishes `done` (a system fence per CTA)
constexpr uint32_t GATHER_TUNE_LOCAL_STORES = 4u;     // measurement only: peers store into their OWN framebuffer (no NVLink traffic; the frame is wrong)
constexpr uint32_t GATHER_TUNE_PERSISTENT_ROOT = 8u;  // the root uses the persistent schedule
constexpr uint32_t GATHER_TUNE_DIRECT_STORES = 32u;   // peers store from the warp tiles directly (32-byte segments) instead of staging whole rows
constexpr uint32_t GATHER_TUNE_INKERNEL_PERSISTENT = 16u;  // peers: persistent schedule, in-kernel `done` (one system fence per resident CTA)

// What the root's completion kernel needs besides the frame: which rows the peers own and the palette to resolve albedo with
struct GatherComplete {
    const uint32_t* done_flags;   // root memory: done[r] word of peer r at done_flags[r * done_stride]
    uint32_t done_stride;         // in u32 words
    uint32_t world, frame_seq;
    uint32_t fill_albedo;         // 1: peers shipped hit id + distance only
    uint32_t width, height, band_shift, band_rotate;
    const uint32_t* hit_id;
    uint32_t* albedo;
    const uint32_t* palette;
    uint32_t n_colors;
    uint64_t timeout_ns;
    uint32_t* error;              // host-mapped word: set to 1 + the late peer's rank on a timeout
};

// Output record of the batched get_by_ray query
struct RayHitRecord {
    uint32_t hit;
    uint32_t palette_value;
    float impact[3];
    float normal[3];
    float distance;
};

struct LaunchConfig {
    int sm_count = 148;
    bool persistent = false;          // warp-granular dynamic tile schedule instead of one CTA per 32x8 block
    bool refill = false;              // persistent schedule whose finished lanes take new pixels (experiment; kernels.cu: render_refill_body)
    bool staged_stores = false;       // static schedule with 32x4-pixel CTAs that write whole 128-byte rows (gather peers: kernels.cu)
    uint32_t* tile_counters = nullptr;  // device, two u32 ticket counters (ping-pong across launches)
};

cudaError_t launch_render(const DeviceTree& tree, const FrameParams& frame, const LaunchConfig& cfg, cudaStream_t stream);
// Gather protocol, device side (multi_gpu.cu drives it):
//   launch_wait_flag      one thread spins (acquire, system scope) until *flag >= want; a peer waits for the root's `go`
//   launch_gather_complete the root waits for every peer's done word, then (8-byte wire format) resolves the albedo of
//                          the peers' rows. Both give up after timeout_ns and report through the host-mapped error word.
cudaError_t launch_wait_flag(const uint32_t* flag, uint32_t want, uint64_t timeout_ns, uint32_t* error, uint32_t error_code,
                             cudaStream_t stream);
cudaError_t launch_gather_complete(const GatherComplete& g, int sm_count, cudaStream_t stream);
//   launch_order_ctas     order[] = the block ids 0..n-1: the classes holding the `head` costliest blocks first, heaviest class
//                         first, then all other blocks in raster order (FrameParams::cta_order)
cudaError_t launch_order_ctas(const uint32_t* cost, uint32_t* order, uint32_t n, uint32_t head, cudaStream_t stream);
//   launch_signal_flag    one thread: system fence, then *flag = value (release, system scope)
cudaError_t launch_signal_flag(uint32_t* flag, uint32_t value, cudaStream_t stream);
cudaError_t launch_rays(const DeviceTree& tree, const float* rays /* [n][6] */, uint64_t n, float viewing_distance,
                        RayHitRecord* out, const LaunchConfig& cfg, cudaStream_t stream);
// Render-data upload, device side: the occupancy bit-bricks (gpu_tree.hpp: brick_bits) of the listed bricks, computed
// from the voxels already resident in `tree.voxels`. A voxel's bit is set unless pix_points_to_empty holds for it
// (reference src/octree/node.rs:405-427): (no colour index or albedo.a == 0) and (no data index or data == 0).
// `tables` (device) = color_words words with bit c set when colour c has albedo.a != 0, then data_words words with bit d
// set when user data d != 0; `handles` = n device-resident brick handles; `bits_out` = the brick_bits array.
cudaError_t launch_occupancy_bits(const DeviceTree& tree, const uint32_t* tables, uint32_t color_words, uint32_t data_words,
                                  const uint32_t* handles, uint32_t n, uint32_t* bits_out, cudaStream_t stream);
// Fills out[0..512) with RAY_TO_NODE mask words, out[512..520) octant masks, then 27*8 step results, evaluated by
// the device closed forms; the host compares them with tables regenerated from the reference's generator logic.
cudaError_t launch_lut_selftest(uint64_t* out /* device, 512 + 8 + 216 entries */, cudaStream_t stream);

}  // namespace svx
