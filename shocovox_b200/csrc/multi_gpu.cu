// Multi-GPU entry points of the C ABI: one frame rendered by several GPUs into ONE framebuffer.
//
// Rays are independent and the tree is read-only while rendering, so every GPU holds a replica of the tree
// and renders the image rows `(row / rows_per_band) % world == rank`. The gather is fused into the viewport kernel:
// a peer's kernel stores its pixels straight into the ROOT GPU's framebuffer (peer-mapped memory: CUDA IPC between
// processes, peer access inside one process), and completion travels through two kinds of flags in the root's memory
// (capi_internal.hpp: GatherSync), all on the device - no host barrier, no collective call, per frame:
//
//   root stream  : viewport kernel (first CTA: go = seq, release.sys; renders the root's rows)
//                  gather_complete_kernel (waits done[r] >= seq for every peer r, acquire.sys; with the 8-byte wire format
//                  it then resolves the albedo of the peers' rows from their hit ids)
//   peer r stream: wait_flag_kernel (go >= seq, polled over NVLink)
//                  viewport kernel (stores into the root's planes; its last CTA: fence.sys, done[r] = seq, release.sys)
//
// `go` is what keeps a peer from overwriting a frame the root's consumer has not finished with: the root publishes it in
// stream order, after everything queued on the root's stream before the render call.
//
// Two ways in: svx_view_gather_* for one process per GPU (torchrun / MPI style; the caller ships the 128-byte handle),
// and svx_multi_* for one process driving all GPUs. The reference has no multi-GPU path; what is kept is
// the contract that the assembled frame equals the single-GPU frame byte for byte.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <vector>

#include "capi_internal.hpp"

using namespace svx;

namespace {

struct HandleBody {  // what svx_gather_handle carries (the struct in the header is the same 128 bytes, field for field)
    cudaIpcMemHandle_t ipc;
    uint32_t width, height, world, rows_per_band;
    int32_t wire;
    int32_t device;
    uint64_t plane_bytes, generation;
    uint8_t reserved[24];
};
static_assert(sizeof(HandleBody) == sizeof(svx_gather_handle), "svx_gather_handle layout");
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");

int32_t check_gather_shape(uint32_t world, uint32_t rows_per_band, int32_t wire) {
    if (world < 1 || world > MAX_GATHER_WORLD) return fail(SVX_E_INVALID_ARGUMENT, "gather: world must be 1..16");
    if (rows_per_band == 0 || (rows_per_band & (rows_per_band - 1)) != 0)
        return fail(SVX_E_INVALID_ARGUMENT, "gather: rows_per_band must be a power of two");
    if (wire != SVX_WIRE_THREE_PLANES && wire != SVX_WIRE_ID_DISTANCE) return fail(SVX_E_INVALID_ARGUMENT, "gather: unknown wire format");
    return SVX_OK;
}

uint32_t tuning_from_env() {
    const char* t = std::getenv("SVX_GATHER_TUNING");
    return t ? (uint32_t)std::strtoul(t, nullptr, 0) : 0u;
}

// Serialises joins and closes of in-process gathers (the root's local_peers list and the peers' local_root pointers);
// taken before any view's own mutex. Render calls never take it.
std::mutex& membership_mu() {
    static std::mutex m;
    return m;
}

// v->mu held (or v is being destroyed): back to a view that renders whole frames into its own framebuffer
void leave_gather_locked(svx_view* v) {
    cudaSetDevice(v->host->device);
    if (v->stream) cudaStreamSynchronize(v->stream);
    if (v->gather_role == GATHER_PEER && v->peer_is_ipc && v->peer_block) cudaIpcCloseMemHandle(v->peer_block);
    v->peer_block = nullptr;
    v->peer_is_ipc = false;
    v->local_root = nullptr;
    v->gather_role = GATHER_NONE;
    v->rank = 0;
    v->world = 1;
    v->frame_seq = 0;
    invalidate_block_order(v);
    if (v->h_error) *v->h_error = 0u;
}

int32_t quiesce(svx_view* v) {
    CUDA_TRY(cudaSetDevice(v->host->device));
    const int32_t drained = retire_locked(v, 0);
    if (drained != SVX_OK) return drained;
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    return SVX_OK;
}

int32_t open_root_locked(svx_view* root, uint32_t world, uint32_t rows_per_band, int32_t wire) {
    const int32_t shape = check_gather_shape(world, rows_per_band, wire);
    if (shape != SVX_OK) return shape;
    if (root->gather_role == GATHER_PEER) return fail(SVX_E_INVALID_ARGUMENT, "gather: this view is a peer of another gather");
    if (root->gather_role == GATHER_ROOT) {
        if (root->world != world || root->band_rows != rows_per_band || root->gather_wire != wire)
            return fail(SVX_E_INVALID_ARGUMENT, "gather: already open with another shape; close it first");
        return SVX_OK;
    }
    if (root->shading || root->compact) return fail(SVX_E_INVALID_ARGUMENT, "gather: not available with the shaded plane or compact rows");
    const int32_t idle = quiesce(root);
    if (idle != SVX_OK) return idle;
    // flags start from zero: every member counts frames from 1
    CUDA_TRY(cudaMemsetAsync(gather_sync_of(root->frame_block, root->plane_bytes), 0, sizeof(GatherSync), root->stream));
    CUDA_TRY(cudaStreamSynchronize(root->stream));
    root->rank = 0;
    root->world = world;
    root->band_rows = rows_per_band;
    root->gather_wire = wire;
    root->frame_seq = 0;
    root->gather_exports = 0;
    root->gather_tuning = tuning_from_env();
    invalidate_block_order(root);
    root->gather_role = GATHER_ROOT;
    return SVX_OK;
}

int32_t become_peer_locked(svx_view* v, uint32_t rank, uint32_t world, uint32_t rows_per_band, int32_t wire, uint32_t width,
                           uint32_t height) {
    if (v->gather_role != GATHER_NONE) return fail(SVX_E_INVALID_ARGUMENT, "gather: this view already belongs to a gather");
    if (rank == 0 || rank >= world) return fail(SVX_E_INVALID_ARGUMENT, "gather: a joining rank must be in 1..world-1 (rank 0 is the root)");
    if (v->width != width || v->height != height) return fail(SVX_E_INVALID_ARGUMENT, "gather: the resolution differs from the root's");
    if (v->shading || v->compact) return fail(SVX_E_INVALID_ARGUMENT, "gather: not available with the shaded plane or compact rows");
    const int32_t idle = quiesce(v);
    if (idle != SVX_OK) return idle;
    CUDA_TRY(cudaMemsetAsync(v->d_cta_counter, 0, sizeof(uint32_t), v->stream));
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    v->rank = rank;
    v->world = world;
    v->band_rows = rows_per_band;
    v->gather_wire = wire;
    v->frame_seq = 0;
    v->gather_tuning = tuning_from_env();
    invalidate_block_order(v);
    return SVX_OK;
}

}  // namespace

extern "C" {

int32_t svx_view_gather_open(svx_view* root, uint32_t world, uint32_t rows_per_band, int32_t wire, svx_gather_handle* out) {
    if (!root) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(root->mu);
    const int32_t opened = open_root_locked(root, world, rows_per_band, wire);
    if (opened != SVX_OK) return opened;
    if (out) {
        HandleBody h{};
        CUDA_TRY(cudaSetDevice(root->host->device));
        CUDA_TRY(cudaIpcGetMemHandle(&h.ipc, root->frame_block));
        h.width = root->width;
        h.height = root->height;
        h.world = world;
        h.rows_per_band = rows_per_band;
        h.wire = wire;
        h.device = root->host->device;
        h.plane_bytes = root->plane_bytes;
        h.generation = root->frame_generation;
        std::memcpy(out, &h, sizeof(h));
        root->gather_exports += 1;
    }
    return SVX_OK;
}

int32_t svx_view_gather_join(svx_view* v, uint32_t rank, const svx_gather_handle* handle) {
    if (!v || !handle) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    HandleBody h;
    std::memcpy(&h, handle, sizeof(h));
    const int32_t shape = check_gather_shape(h.world, h.rows_per_band, h.wire);
    if (shape != SVX_OK) return shape;
    const size_t n = (size_t)h.width * h.height;
    if (n == 0 || h.plane_bytes < n * 4) return fail(SVX_E_INVALID_ARGUMENT, "gather: malformed handle");
    std::lock_guard<std::mutex> lock(v->mu);
    const int32_t ready = become_peer_locked(v, rank, h.world, h.rows_per_band, h.wire, h.width, h.height);
    if (ready != SVX_OK) return ready;
    void* mapped = nullptr;
    const cudaError_t opened = cudaIpcOpenMemHandle(&mapped, h.ipc, cudaIpcMemLazyEnablePeerAccess);
    if (opened != cudaSuccess) {  // not a member after all: back to a whole-frame view
        v->rank = 0;
        v->world = 1;
        return cuda_fail(opened, "cudaIpcOpenMemHandle (the root must live in another process on a GPU with peer access)");
    }
    v->peer_block = mapped;
    v->peer_plane_bytes = (size_t)h.plane_bytes;
    v->peer_is_ipc = true;
    v->local_root = nullptr;
    v->gather_role = GATHER_PEER;
    return SVX_OK;
}

int32_t svx_view_gather_join_local(svx_view* v, uint32_t rank, svx_view* root) {
    if (!v || !root || v == root) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> membership(membership_mu());
    std::lock_guard<std::mutex> lock_root(root->mu);
    std::lock_guard<std::mutex> lock(v->mu);
    if (root->gather_role != GATHER_ROOT) return fail(SVX_E_INVALID_ARGUMENT, "gather: open the root first (svx_view_gather_open)");
    const int32_t ready = become_peer_locked(v, rank, root->world, root->band_rows, root->gather_wire, root->width, root->height);
    if (ready != SVX_OK) return ready;
    const int mine = v->host->device, theirs = root->host->device;
    if (mine != theirs) {
        int can = 0;
        cudaError_t e = cudaDeviceCanAccessPeer(&can, mine, theirs);
        if (e == cudaSuccess && can) {
            e = cudaSetDevice(mine);
            if (e == cudaSuccess) e = cudaDeviceEnablePeerAccess(theirs, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                e = cudaSuccess;
            }
        }
        if (e != cudaSuccess || !can) {
            v->rank = 0;
            v->world = 1;
            return e != cudaSuccess ? cuda_fail(e, "peer access to the root's device") : fail(SVX_E_CUDA, "gather: no peer access between the two devices");
        }
    }
    v->peer_block = root->frame_block;
    v->peer_plane_bytes = root->plane_bytes;
    v->peer_is_ipc = false;
    v->local_root = root;
    v->gather_role = GATHER_PEER;
    root->gather_exports += 1;
    root->local_peers.push_back(v);
    return SVX_OK;
}

int32_t svx_view_gather_close(svx_view* v) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> membership(membership_mu());
    {
        std::lock_guard<std::mutex> lock(v->mu);
        if (v->gather_role == GATHER_NONE) return SVX_OK;
        if (v->local_root) {  // a peer of a root in this process: the root forgets it
            std::vector<svx_view*>& peers = v->local_root->local_peers;
            peers.erase(std::remove(peers.begin(), peers.end(), v), peers.end());
        }
        leave_gather_locked(v);
    }
    // a root takes its in-process peers with it: they store through this view's own device pointers, which the caller
    // may free next (peers in other processes hold CUDA IPC mappings and must be closed by their owners first)
    std::vector<svx_view*> peers;
    peers.swap(v->local_peers);
    for (svx_view* p : peers) {
        std::lock_guard<std::mutex> lock(p->mu);
        leave_gather_locked(p);
    }
    debug_stale("svx_view_gather_close: exit");
    return SVX_OK;
}

int32_t svx_view_gather_info(const svx_view* v, int32_t* role, uint32_t* rank, uint32_t* world, uint32_t* frames) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    if (role) *role = v->gather_role;
    if (rank) *rank = v->rank;
    if (world) *world = v->world;
    if (frames) *frames = v->frame_seq;
    return SVX_OK;
}

}  // extern "C"

// ---- svx_multi: one process, all GPUs ----------------------------------------------------------------------------------
struct svx_multi {
    const svx_octree* tree = nullptr;
    uint32_t n = 0, width = 0, height = 0, band = 8;
    int32_t wire = SVX_WIRE_THREE_PLANES;
    std::vector<int32_t> devices;
    std::vector<svx_gpu_host*> hosts;
    std::vector<svx_view*> gather;  // [0] = root on devices[0]; the others store into it
    std::vector<svx_view*> local;   // tile shards with their own framebuffers (read-back over every GPU's own PCIe link); lazy
    std::vector<svx_view*> whole;   // unsharded views for pose batches; lazy
    svx_viewport viewport{};
    int32_t glass_mode = SVX_GLASS_AT_FOV;
    float viewing_distance = 3.402823466e+38f;
    std::mutex mu;
};

namespace {

int32_t configure(svx_multi* m, svx_view* v) {
    int32_t s = svx_view_set_glass_mode(v, m->glass_mode);
    if (s == SVX_OK) s = svx_view_set_viewing_distance(v, m->viewing_distance);
    return s;
}

int32_t ensure_views(svx_multi* m, std::vector<svx_view*>* set, bool sharded) {
    if (!set->empty()) return SVX_OK;
    for (uint32_t i = 0; i < m->n; ++i) {
        svx_view* v = nullptr;
        int32_t s = svx_gpu_host_create_view(m->hosts[i], 0, &m->viewport, m->width, m->height, &v);
        if (s == SVX_OK) s = configure(m, v);
        if (s == SVX_OK && sharded) s = svx_view_set_shard(v, i, m->n, m->band);
        if (s != SVX_OK) {
            svx_view_free(v);
            for (svx_view* w : *set) svx_view_free(w);
            set->clear();
            return s;
        }
        set->push_back(v);
    }
    return SVX_OK;
}

template <typename F>
int32_t for_all_views(svx_multi* m, F&& f) {
    for (auto* set : {&m->gather, &m->local, &m->whole})
        for (svx_view* v : *set) {
            const int32_t s = f(v);
            if (s != SVX_OK) return s;
        }
    return SVX_OK;
}

}  // namespace

extern "C" {

void svx_multi_free(svx_multi* m) {
    if (!m) return;
    // peers first: they hold pointers into the root's frame
    for (size_t i = m->gather.size(); i-- > 0;) svx_view_free(m->gather[i]);
    for (svx_view* v : m->local) svx_view_free(v);
    for (svx_view* v : m->whole) svx_view_free(v);
    for (svx_gpu_host* h : m->hosts) svx_gpu_host_free(h);
    octree_release(m->tree);
    delete m;
}

int32_t svx_multi_create(const svx_octree* tree, const int32_t* devices, uint32_t n, const svx_viewport* viewport, uint32_t width,
                         uint32_t height, uint32_t rows_per_band, int32_t wire, svx_multi** out) {
    if (!tree || !devices || !viewport || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    debug_stale("svx_multi_create: entry");
    const int32_t shape = check_gather_shape(n, rows_per_band, wire);
    if (shape != SVX_OK) return shape;
    svx_multi* m = new (std::nothrow) svx_multi();
    if (!m) return fail(SVX_E_OUT_OF_MEMORY, "allocation failed");
    m->tree = tree;
    octree_retain(tree);
    m->n = n;
    m->width = width;
    m->height = height;
    m->band = rows_per_band;
    m->wire = wire;
    m->viewport = *viewport;
    m->devices.assign(devices, devices + n);
    int32_t s = SVX_OK;
    for (uint32_t i = 0; i < n && s == SVX_OK; ++i) {  // one replica of the tree per GPU (a device may be listed twice: tests)
        svx_gpu_host* h = nullptr;
        s = svx_gpu_host_create(tree, devices[i], &h);
        if (s == SVX_OK) m->hosts.push_back(h);
    }
    for (uint32_t i = 0; i < n && s == SVX_OK; ++i) {
        svx_view* v = nullptr;
        s = svx_gpu_host_create_view(m->hosts[i], 0, viewport, width, height, &v);
        if (s == SVX_OK) m->gather.push_back(v);
    }
    if (s == SVX_OK) s = svx_view_gather_open(m->gather[0], n, rows_per_band, wire, nullptr);
    for (uint32_t i = 1; i < n && s == SVX_OK; ++i) s = svx_view_gather_join_local(m->gather[i], i, m->gather[0]);
    if (s != SVX_OK) {
        svx_multi_free(m);
        return s;
    }
    *out = m;
    return SVX_OK;
}

uint32_t svx_multi_device_count(const svx_multi* m) { return m ? m->n : 0; }
svx_view* svx_multi_view(svx_multi* m, uint32_t i) { return (m && i < m->n) ? m->gather[i] : nullptr; }

int32_t svx_multi_set_viewport(svx_multi* m, const svx_viewport* vp) {
    if (!m || !vp) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    const int32_t valid = validate_viewport(*vp);
    if (valid != SVX_OK) return valid;
    std::lock_guard<std::mutex> lock(m->mu);
    m->viewport = *vp;
    return SVX_OK;  // handed to the views by the render calls
}
int32_t svx_multi_set_glass_mode(svx_multi* m, int32_t mode) {
    if (!m || (mode != SVX_GLASS_AT_FOV && mode != SVX_GLASS_AT_FRUSTUM_Z)) return fail(SVX_E_INVALID_ARGUMENT, "bad mode");
    std::lock_guard<std::mutex> lock(m->mu);
    m->glass_mode = mode;
    return for_all_views(m, [&](svx_view* v) { return svx_view_set_glass_mode(v, mode); });
}
int32_t svx_multi_set_viewing_distance(svx_multi* m, float viewing_distance) {
    if (!m) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(m->mu);
    m->viewing_distance = viewing_distance;
    return for_all_views(m, [&](svx_view* v) { return svx_view_set_viewing_distance(v, viewing_distance); });
}
int32_t svx_multi_reload(svx_multi* m) {
    if (!m) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(m->mu);
    for (svx_gpu_host* h : m->hosts) {
        const int32_t s = svx_gpu_host_reload(h);
        if (s != SVX_OK) return s;
    }
    return SVX_OK;
}

// One frame, assembled in the framebuffer of devices[0]. Peers are queued first (their kernels wait on the device for
// the root's go flag), the root last, so all GPUs start within a flag's latency of each other.
int32_t svx_multi_render(svx_multi* m, svx_frame* out) {
    if (!m) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(m->mu);
    for (uint32_t k = 0; k < m->n; ++k) {
        svx_view* v = m->gather[m->n - 1 - k];
        int32_t s = svx_view_set_viewport(v, &m->viewport);
        if (s != SVX_OK) return s;
        std::lock_guard<std::mutex> view_lock(v->mu);
        CUDA_TRY(cudaSetDevice(v->host->device));
        s = render_locked(v, out != nullptr && k == m->n - 1);  // the root's event pair is the frame time reported below
        if (s != SVX_OK) return s;
    }
    if (!out) return SVX_OK;
    svx_view* root = m->gather[0];
    std::lock_guard<std::mutex> root_lock(root->mu);
    CUDA_TRY(cudaSetDevice(root->host->device));
    CUDA_TRY(cudaStreamSynchronize(root->stream));
    const int32_t arrived = check_view_error(root);
    if (arrived != SVX_OK) return arrived;
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, root->ev_start, root->ev_stop));
    out->width = m->width;
    out->height = m->height;
    out->row_begin = 0;
    out->row_end = m->height;
    out->hit_id = root->d_hit_id;
    out->albedo = root->d_albedo;
    out->distance = root->d_distance;
    out->kernel_ms = ms;  // root's viewport kernel + the wait for the slowest peer: the frame time
    return SVX_OK;
}

// The same frame delivered into HOST planes. No NVLink gather here: every GPU renders its rows into its own framebuffer
// and copies exactly those rows into the host planes over its own PCIe link - the frame is assembled in host memory at the
// aggregate device->host bandwidth of all GPUs. Pinned (page-locked, portable) host buffers let the copies overlap.
int32_t svx_multi_render_to_host(svx_multi* m, uint32_t* hit_id, uint32_t* albedo, float* distance) {
    if (!m) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(m->mu);
    int32_t s = ensure_views(m, &m->local, true);
    if (s != SVX_OK) return s;
    for (svx_view* v : m->local) {
        s = svx_view_set_viewport(v, &m->viewport);
        if (s != SVX_OK) return s;
        std::lock_guard<std::mutex> view_lock(v->mu);
        CUDA_TRY(cudaSetDevice(v->host->device));
        s = render_locked(v, false);
        if (s == SVX_OK) s = copy_frame_to_host(v, v->stream, hit_id, albedo, distance);
        if (s != SVX_OK) return s;
    }
    for (svx_view* v : m->local) {
        s = svx_view_synchronize(v);
        if (s != SVX_OK) return s;
    }
    return SVX_OK;
}

// Batch mode (BASELINE config 5): pose k is rendered by GPU k % n into host planes [n_poses][h*w]; no exchange at all.
int32_t svx_multi_render_poses(svx_multi* m, const svx_viewport* poses, uint32_t n_poses, uint32_t* hit_id, uint32_t* albedo,
                               float* distance, float* kernel_ms_total) {
    if (!m || (n_poses && !poses)) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    for (uint32_t k = 0; k < n_poses; ++k) {
        const int32_t valid = validate_viewport(poses[k]);
        if (valid != SVX_OK) return valid;
    }
    std::lock_guard<std::mutex> lock(m->mu);
    int32_t s = ensure_views(m, &m->whole, false);
    if (s != SVX_OK) return s;
    const size_t px = (size_t)m->width * m->height;
    // round-robin submission keeps every GPU's two-frame pipeline full (svx_view_render_to_host_async)
    for (uint32_t k = 0; k < n_poses; ++k) {
        svx_view* v = m->whole[k % m->n];
        s = svx_view_set_viewport(v, &poses[k]);
        if (s == SVX_OK)
            s = svx_view_render_to_host_async(v, hit_id ? hit_id + k * px : nullptr, albedo ? albedo + k * px : nullptr,
                                              distance ? distance + k * px : nullptr);
        if (s != SVX_OK) return s;
    }
    float total = 0.0f;
    for (svx_view* v : m->whole) {
        float ms = 0.0f;
        s = svx_view_wait_host(v, 0, &ms);
        if (s != SVX_OK) return s;
        total += ms;
    }
    if (kernel_ms_total) *kernel_ms_total = total;
    return SVX_OK;
}

}  // extern "C"
