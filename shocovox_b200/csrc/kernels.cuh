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
    uint32_t* go_flag;
    uint32_t* done_flag;
    uint32_t* cta_counter;
    uint32_t frame_seq;
    uint32_t gather_tuning;  // GATHER_TUNE_* bits (multi_gpu.cu), 0 = the defaults
};

// Tuning bits of the gather (SVX_GATHER_TUNING, read when a view opens / joins a gather). Default 0 = what measured best
// (profiles/r02_gather_probe_*.json): peers run the persistent schedule, so the system-scope fence that must precede the
// `done` flag is paid once per resident CTA instead of once per 16x8-pixel block - a membar.sys per block cost 18 % of the
// peer's kernel even with local stores.
constexpr uint32_t GATHER_TUNE_CTA_FENCE_GPU = 1u;  // retiring CTAs fence at gpu scope, only the publishing one at system scope
constexpr uint32_t GATHER_TUNE_STATIC_PEERS = 2u;   // peers use the static schedule (one CTA and one fence per pixel block)
constexpr uint32_t GATHER_TUNE_LOCAL_STORES = 4u;   // measurement only: peers store into their OWN framebuffer (no NVLink traffic; the frame is wrong)
constexpr uint32_t GATHER_TUNE_PERSISTENT_ROOT = 8u;  // the root uses the persistent schedule as well
constexpr uint32_t GATHER_TUNE_SIGNAL_KERNEL = 16u;   // peers: static schedule, no fence in the viewport kernel; a one-thread kernel queued
                                                      // behind it publishes `done` (kernel completion orders the stores)

// What the root's completion kernel needs besides the frame: which rows the peers own and the palette to resolve albedo with
struct GatherComplete {
    const uint32_t* done_flags;   // root memory: done[r] word of peer r at done_flags[r * done_stride]
    uint32_t done_stride;         // in u32 words
    uint32_t world, frame_seq;
    uint32_t fill_albedo;         // 1: peers shipped hit id + distance only
    uint32_t width, height, band_shift;
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
