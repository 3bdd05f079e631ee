// Handle types behind the C ABI (include/shocovox_b200.h) and the helpers capi.cu and multi_gpu.cu share.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>
#include <vector>
#include <shared_mutex>
#include <string>

#include "../../include/shocovox_b200.h"
#include "gpu_tree.hpp"
#include "host_octree.hpp"
#include "kernels.cuh"

namespace svx {

void debug_stale(const char* where);  // SVX_DEBUG_CUDA=1: report and clear a left-over CUDA error at the named point
int32_t fail(int32_t code, const std::string& msg);
int32_t cuda_fail(cudaError_t e, const char* what);
#define CUDA_TRY(expr)                                             \
    do {                                                           \
        cudaError_t e__ = (expr);                                  \
        if (e__ != cudaSuccess) return svx::cuda_fail(e__, #expr); \
    } while (0)

// Flags of one tile-sharded gather. They live at the end of the ROOT view's frame allocation, so one CUDA IPC handle
// exports framebuffer and flags together. Every word other GPUs write or poll sits in its own 128-byte line.
constexpr uint32_t MAX_GATHER_WORLD = 16;
struct GatherSync {
    uint32_t go;  // written by the root's viewport kernel, polled by the peers
    uint32_t pad0[31];
    struct {
        uint32_t seq;  // written by peer r's viewport kernel (its last CTA), polled by the root
        uint32_t pad[31];
    } done[MAX_GATHER_WORLD];
};
static_assert(sizeof(GatherSync) == 128 * (1 + MAX_GATHER_WORLD), "one line per flag");

enum GatherRole : int32_t { GATHER_NONE = 0, GATHER_ROOT = 1, GATHER_PEER = 2 };

}  // namespace svx

struct svx_gpu_host;
// Lifetimes: a host keeps its octree alive and a view keeps its host alive (intrusive counts), so the *_free calls may come
// in any order - language bindings whose finalisers run in no particular order (Python's cycle collector, Rust drop order
// of unrelated owners) cannot leave a dangling handle behind. A freed handle must still not be USED by the caller.
struct svx_octree {
    mutable std::atomic<int> refs{1};  // the caller's handle + one per host / multi built on it
    svx::HostOctree* tree = nullptr;
    // svx_octree_get_by_ray[_at_lod]: a device copy created on first use (device 0) and reloaded before every query
    svx_gpu_host* ray_host = nullptr;
    std::mutex ray_mu;
};

struct svx_gpu_host {
    std::atomic<int> refs{1};          // the caller's handle + one per view
    bool holds_octree_ref = true;      // false for the octree's own ray_host (it is destroyed WITH the octree)
    const svx_octree* octree = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // `dev` and the arrays behind it: launches of the views take it shared around "snapshot dev + launch", a reload takes
    // it exclusively, drains the device and only then replaces / frees arrays (lock order: view->mu or host->mu, then dev_mu)
    std::shared_mutex dev_mu;
    svx_gpu_stats stats{};
    svx::DeviceTree dev{};
    void* d_node_rec = nullptr;  // 64-byte node records: head | 8 slots | bounds
    void* d_node_mip = nullptr;
    void* d_voxels = nullptr;
    void* d_brick_bits = nullptr;
    void* d_palette = nullptr;
    void* d_ray_lut = nullptr;  // RAY_TO_NODE_OCCUPANCY_BITMASK_LUT as [dir][cell] {lo, hi}, regenerated from the generator logic
    void* d_data_palette = nullptr;  // bit tables "colour shows" / "data carries", input of the occupancy-bit kernel
    void* d_handles = nullptr;       // brick handles of the current upload, input of the occupancy-bit kernel
    size_t data_palette_capacity = 0, handle_capacity = 0;
    size_t node_capacity = 0;     // nodes the node_head / node_slot allocations hold
    size_t palette_capacity = 0;  // colours the palette allocation holds
    size_t brick_capacity = 0;    // bricks the voxels / brick_bits allocations hold
    svx::LaunchConfig cfg;
    uint64_t launches = 0;
    uint64_t uploaded_revision = ~0ull;
    bool uploaded = false;
    svx_upload_stats last_upload{};
    // scratch for get_by_rays
    float* d_rays = nullptr;
    svx::RayHitRecord* d_hits = nullptr;
    uint64_t ray_capacity = 0;
};

struct svx_view {
    svx_gpu_host* host = nullptr;
    svx_viewport viewport{};
    int32_t glass_mode = SVX_GLASS_AT_FOV;
    float viewing_distance = 3.402823466e+38f;  // f32::MAX: Octree::get_by_ray (raytracing_on_cpu.rs:316-318)
    uint32_t width = 0, height = 0;
    uint32_t rank = 0, world = 1, band_rows = 8, compact = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t tm_start = nullptr, tm_stop = nullptr;
    uint32_t* d_counters = nullptr;  // two ticket counters of the persistent schedule (ping-pong across launches)
    uint32_t counter_slot = 0;
    bool persistent = false;
    bool refill = false;  // with `persistent`: the lane-refill schedule (svx_view_set_schedule(view, 2))
    // SVX_REFILL_STEPS / SVX_REFILL_MIN_IDLE / SVX_REFILL_UNIT at view creation (tools/refill_probe.py)
    uint32_t refill_steps = 64, refill_min_idle = 16, refill_unit_tiles = 1;  // the best of the sweep (profiles/r02_experiment_lane_refill.json)
    // heaviest-first block order of the static schedule (kernels.cuh: FrameParams::cta_order): cost and order arrays of
    // `order_ctas` blocks; `order_valid` once a frame has been recorded and sorted for the current resolution / shard
    uint32_t* d_cta_cost = nullptr;
    uint32_t* d_cta_order = nullptr;
    uint32_t order_ctas = 0;
    bool order_valid = false;
    uint32_t order_head_pct = 100;  // SVX_CTA_ORDER_HEAD_PCT: share of the blocks dispatched heaviest-first; the rest keeps raster order
    int32_t order_policy = 0;  // SVX_CTA_ORDER: 0 never (default), 1 for the shards of a frame split 4+ ways, 2 always
    void* d_flush = nullptr;
    size_t flush_bytes = 0;
    // Framebuffer: ONE allocation = hit_id | albedo | distance | GatherSync, the planes `plane_bytes` apart
    void* frame_block = nullptr;
    size_t plane_bytes = 0;
    uint64_t frame_generation = 0;  // bumped whenever frame_block is reallocated
    uint32_t* d_hit_id = nullptr;
    uint32_t* d_albedo = nullptr;
    float* d_distance = nullptr;
    // optional shaded plane (svx_view_set_shading): the pixel of the reference's caller loops, single-buffered
    uint32_t* d_shaded = nullptr;
    bool shading = false;
    float light[3] = {0.0f, 0.0f, 0.0f};
    uint64_t launches = 0;
    // Tile-sharded gather (multi_gpu.cu). Root: `world` GPUs render into this view's framebuffer. Peer: this view's
    // kernel stores into the root's framebuffer (`peer_block`: a CUDA IPC mapping, or the root's own pointer when the root
    // lives in this process) and announces completion in the root's GatherSync.
    int32_t gather_role = svx::GATHER_NONE;
    int32_t gather_wire = SVX_WIRE_THREE_PLANES;
    uint32_t frame_seq = 0;     // frames rendered in this gather; every member counts the same sequence
    uint32_t gather_exports = 0;  // root: handles handed out / local peers attached
    void* peer_block = nullptr;
    size_t peer_plane_bytes = 0;
    bool peer_is_ipc = false;
    svx_view* local_root = nullptr;
    // root: the peers of this process that store through this view's own pointers (svx_view_gather_join_local). Closing or
    // freeing the root detaches them first, so none is left pointing into a freed frame. Guarded by multi_gpu.cu's
    // membership mutex, not by `mu`.
    std::vector<svx_view*> local_peers;
    uint32_t* d_cta_counter = nullptr;  // peer: retired CTAs of the launch in flight
    uint32_t* h_error = nullptr;        // host-mapped word the wait kernels write on a timeout (0 = fine)
    uint64_t gather_timeout_ns = 5000000000ull;
    uint32_t gather_tuning = 0;  // GATHER_TUNE_* (kernels.cuh), from SVX_GATHER_TUNING at open / join
    // Pipelined read-back (svx_view_render_to_host_async): two framebuffer slots (slot 0 = the planes above, slot 1 =
    // alt_*), kernels on `stream`, device->host copies on `copy_stream`, so frame i's copy overlaps frame i+1's kernel.
    cudaStream_t copy_stream = nullptr;
    uint32_t* alt_hit_id = nullptr;
    uint32_t* alt_albedo = nullptr;
    float* alt_distance = nullptr;
    cudaEvent_t slot_start[2] = {nullptr, nullptr}, slot_rendered[2] = {nullptr, nullptr}, slot_copied[2] = {nullptr, nullptr};
    bool slot_busy[2] = {false, false};
    uint64_t async_frames = 0;   // frames submitted through the pipelined path
    float async_kernel_ms = 0.0f;  // summed kernel time of the pipelined frames retired so far
    uint32_t target_slot = 0;    // which slot make_frame_constants points the kernel at
    std::mutex mu;
};

namespace svx {

inline GatherSync* gather_sync_of(void* frame_block, size_t plane_bytes) {
    return reinterpret_cast<GatherSync*>(static_cast<char*>(frame_block) + 3 * plane_bytes);
}

// capi.cu
void octree_retain(const svx_octree* t);
void octree_release(const svx_octree* t);
void host_release(svx_gpu_host* h);
int32_t validate_viewport(const svx_viewport& vp);
// one frame on the view's stream (gather roles included); v->mu held. timed: record ev_start / ev_stop around it
int32_t render_locked(svx_view* v, bool timed);
int32_t retire_locked(svx_view* v, uint32_t keep);
int32_t check_view_error(svx_view* v);    // after a synchronise: did a gather wait give up?
void invalidate_block_order(svx_view* v); // resolution / shard changed: the recorded block costs no longer apply
// device -> host copies of the rows this view owns (all of them unless it is a local shard), on `stream`
int32_t copy_frame_to_host(svx_view* v, cudaStream_t stream, uint32_t* hit_id, uint32_t* albedo, float* distance);

}  // namespace svx
