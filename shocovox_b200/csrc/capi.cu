// C ABI of shocovox_b200 (include/shocovox_b200.h): octree construction on the host, render-data upload,
// viewport rendering and batched ray queries on the GPU. There is no CPU ray path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "capi_internal.hpp"
#include "vox_import.hpp"

using namespace svx;

#ifndef SVX_DEFAULT_PERSISTENT
#define SVX_DEFAULT_PERSISTENT false
#endif

namespace {
thread_local std::string g_last_error;
}
namespace svx {
// SVX_DEBUG_CUDA=1: report (and clear) an error some earlier runtime call left behind, at the named point
void debug_stale(const char* where) {
    static const bool on = std::getenv("SVX_DEBUG_CUDA") != nullptr;
    if (!on) return;
    const cudaError_t e = cudaGetLastError();
    int dev = -1;
    cudaGetDevice(&dev);
    std::fprintf(stderr, "[svx] %s: last error %s, current device %d\n", where, cudaGetErrorName(e), dev);
}
int32_t fail(int32_t code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int32_t cuda_fail(cudaError_t e, const char* what) {
    return fail(SVX_E_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
}  // namespace svx

namespace {

// ---- the reference's tables, regenerated from their generator logic (reference src/spatial/lut.rs:12-152), used
// ---- ONLY to validate the device closed forms once per process
void host_tables(uint64_t* ray2node /*512*/, uint64_t* octmask /*8*/, uint32_t* step /*216*/) {
    auto offs = [](int o, int axis) { return axis == 0 ? (o & 1) : (axis == 1 ? ((o >> 2) & 1) : ((o >> 1) & 1)); };
    for (int o = 0; o < 8; ++o) {
        uint64_t m = 0;
        for (int x = 2 * offs(o, 0); x < 2 * offs(o, 0) + 2; ++x)
            for (int y = 2 * offs(o, 1); y < 2 * offs(o, 1) + 2; ++y)
                for (int z = 2 * offs(o, 2); z < 2 * offs(o, 2) + 2; ++z) m |= occupancy_box(x, y, z, 1, 4);
        octmask[o] = m;
    }
    for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y)
            for (int z = 0; z < 4; ++z)
                for (int dx = -1; dx <= 1; dx += 2)
                    for (int dy = -1; dy <= 1; dy += 2)
                        for (int dz = -1; dz <= 1; dz += 2) {
                            const int dir = ((1.0f + (float)dx) >= 1.0f) + 2 * ((1.0f + (float)dz) >= 1.0f) +
                                            4 * ((1.0f + (float)dy) >= 1.0f);
                            const int mx = std::clamp(x + dx * 4, 0, 3), my = std::clamp(y + dy * 4, 0, 3),
                                      mz = std::clamp(z + dz * 4, 0, 3);
                            uint64_t m = 0;
                            for (int bx = std::min(mx, x); bx <= std::max(mx, x); ++bx)
                                for (int by = std::min(my, y); by <= std::max(my, y); ++by)
                                    for (int bz = std::min(mz, z); bz <= std::max(mz, z); ++bz)
                                        m |= occupancy_box(bx, by, bz, 1, 4);
                            ray2node[(x + 4 * y + 16 * z) * 8 + dir] = m;
                        }
    for (int o = 0; o < 8; ++o)
        for (int sx = -1; sx <= 1; ++sx)
            for (int sy = -1; sy <= 1; ++sy)
                for (int sz = -1; sz <= 1; ++sz) {
                    const float c[3] = {3.0f + offs(o, 0) * 6.0f + sx * 6.0f, 3.0f + offs(o, 1) * 6.0f + sy * 6.0f,
                                        3.0f + offs(o, 2) * 6.0f + sz * 6.0f};
                    uint32_t res;
                    if (c[0] < 0 || c[0] > 12 || c[1] < 0 || c[1] > 12 || c[2] < 0 || c[2] > 12)
                        res = 8;
                    else
                        res = (c[0] >= 6.0f) + 2 * (c[2] >= 6.0f) + 4 * (c[1] >= 6.0f);
                    step[((sx + 1) * 9 + (sy + 1) * 3 + (sz + 1)) * 8 + o] = res;
                }
}

// RAY_TO_NODE_OCCUPANCY_BITMASK_LUT in the layout the kernels read (DeviceTree::ray_lut): [direction octant][cell] {lo, hi}
void make_ray_lut(uint32_t* lut /* 8 * 64 * 2 */) {
    uint64_t r2n[512], om[8];
    uint32_t st[216];
    host_tables(r2n, om, st);
    for (int cell = 0; cell < 64; ++cell)
        for (int dir = 0; dir < 8; ++dir) {
            lut[(dir * 64 + cell) * 2] = (uint32_t)r2n[cell * 8 + dir];
            lut[(dir * 64 + cell) * 2 + 1] = (uint32_t)(r2n[cell * 8 + dir] >> 32);
        }
}

std::once_flag g_selftest_once;
int32_t g_selftest_status = SVX_OK;
std::string g_selftest_error;

void run_selftest(cudaStream_t stream) {
    uint64_t* dev = nullptr;
    std::vector<uint64_t> got(736);
    cudaError_t e = cudaMalloc(&dev, got.size() * 8);
    if (e == cudaSuccess) e = launch_lut_selftest(dev, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(got.data(), dev, got.size() * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (dev) cudaFree(dev);
    if (e != cudaSuccess) {
        g_selftest_status = SVX_E_CUDA;
        g_selftest_error = std::string("LUT self-test launch failed: ") + cudaGetErrorString(e);
        return;
    }
    uint64_t r2n[512], om[8];
    uint32_t st[216];
    host_tables(r2n, om, st);
    for (int i = 0; i < 512; ++i)
        if (got[i] != r2n[i]) g_selftest_status = SVX_E_CUDA;
    for (int i = 0; i < 8; ++i)
        if (got[512 + i] != om[i]) g_selftest_status = SVX_E_CUDA;
    for (int i = 0; i < 216; ++i)
        if (got[520 + i] != st[i]) g_selftest_status = SVX_E_CUDA;
    if (g_selftest_status != SVX_OK) g_selftest_error = "device closed forms disagree with the reference look-up tables";
}

struct Vec3 {
    float x, y, z;
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }

}  // namespace

namespace {

void free_device_tree(svx_gpu_host* h) {
    cudaFree(h->d_node_rec);
    cudaFree(h->d_node_mip);
    cudaFree(h->d_voxels);
    cudaFree(h->d_brick_bits);
    cudaFree(h->d_palette);
    cudaFree(h->d_data_palette);
    cudaFree(h->d_handles);
    h->d_node_mip = nullptr;
    h->d_node_rec = h->d_voxels = h->d_brick_bits = h->d_palette = h->d_data_palette = h->d_handles = nullptr;
    h->node_capacity = h->palette_capacity = h->brick_capacity = h->data_palette_capacity = h->handle_capacity = 0;
    h->uploaded = false;
}

// Makes `*ptr` hold at least `need` elements of `elem` bytes (capacity grows by 1.5x), optionally keeping the first
// `keep` elements (device-to-device copy on the host's stream).
cudaError_t grow_device_array(void** ptr, size_t* capacity, size_t need, size_t elem, size_t keep, cudaStream_t stream) {
    if (need <= *capacity && *ptr) return cudaSuccess;
    const size_t cap = std::max<size_t>(std::max(need, *capacity + *capacity / 2), 1);
    void* fresh = nullptr;
    cudaError_t e = cudaMalloc(&fresh, std::max<size_t>(cap * elem, 16));
    if (e != cudaSuccess) return e;
    if (*ptr && keep) {
        e = cudaMemcpyAsync(fresh, *ptr, keep * elem, cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            cudaFree(fresh);
            return e;
        }
    }
    cudaFree(*ptr);
    *ptr = fresh;
    *capacity = cap;
    return cudaSuccess;
}

// interleave head / slots / bounds into the 64-byte records (gpu_tree.hpp: DeviceTree::node_rec)
void pack_node_records(const SerialisedNodes& s, uint32_t* records) {
    for (size_t i = 0; i < s.node_head.size(); ++i) {
        uint32_t* rec = records + i * 16;
        std::memcpy(rec, &s.node_head[i], 16);
        std::memcpy(rec + 4, s.node_slot.data() + i * 8, 32);
        std::memcpy(rec + 12, s.node_bounds.data() + i * 4, 16);
    }
}

// Render-data upload (OctreeGPUHost::create_new_view / write_to_gpu, src/raytracing/bevy/data.rs:111, :365).
// The node tables and the palette are re-serialised and replaced; bricks are mirrored by pool handle, and only the
// ones written since the revision this host last uploaded are copied (all of them the first time).
int32_t upload(svx_gpu_host* h) {
    const HostOctree& tree = *h->octree->tree;
    SerialisedNodes s;
    serialise_nodes(tree, &s);
    const size_t pool = tree.brick_pool_size();
    const size_t vol = tree.brick_volume(), words = s.bit_words;
    // the brick DDA addresses brick_bits with a 32-bit word offset (traverse.cuh: traverse_brick); 2^32 words are
    // 2^37 voxels, far beyond what the u32 voxel array of the same tree could hold in 180 GB
    if (pool * words > 0xFFFFFFFFull) return fail(SVX_E_OUT_OF_MEMORY, "brick pool exceeds 2^32 occupancy words");
    const bool first = !h->uploaded;
    svx_upload_stats up{};
    up.full = first ? 1u : 0u;

    // nodes + palette: small, replaced wholesale
    const size_t n_nodes = s.node_head.size();
    size_t head_capacity = h->node_capacity, mip_capacity = h->node_capacity;
    CUDA_TRY(grow_device_array(&h->d_node_rec, &head_capacity, n_nodes, 64, 0, h->stream));
    CUDA_TRY(grow_device_array(&h->d_node_mip, &mip_capacity, head_capacity, 4, 0, h->stream));
    h->node_capacity = head_capacity;
    CUDA_TRY(grow_device_array(&h->d_palette, &h->palette_capacity, s.palette.size(), 4, 0, h->stream));
    std::vector<uint32_t> records(n_nodes * 16);
    pack_node_records(s, records.data());
    CUDA_TRY(cudaMemcpyAsync(h->d_node_rec, records.data(), records.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_node_mip, s.node_mip.data(), n_nodes * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));  // `records` is a local staging buffer
    CUDA_TRY(cudaMemcpyAsync(h->d_palette, s.palette.data(), s.palette.size() * 4, cudaMemcpyHostToDevice, h->stream));
    up.bytes += n_nodes * (sizeof(NodeHead) + 52) + s.palette.size() * 4;

    // bricks: grow keeping what is resident, then copy the runs of handles written since the last upload
    const size_t resident = first ? 0 : std::min(h->brick_capacity, pool);
    size_t voxel_capacity = h->brick_capacity, bits_capacity = h->brick_capacity;
    CUDA_TRY(grow_device_array(&h->d_voxels, &voxel_capacity, pool, vol * 4, resident, h->stream));
    CUDA_TRY(grow_device_array(&h->d_brick_bits, &bits_capacity, voxel_capacity, words * 4, resident, h->stream));
    h->brick_capacity = voxel_capacity;
    std::vector<uint32_t> handles;
    for (size_t a = 0; a < pool;) {
        if (!first && tree.brick_revision((uint32_t)a) <= h->uploaded_revision) {
            ++a;
            continue;
        }
        size_t b = a + 1;
        while (b < pool && (first || tree.brick_revision((uint32_t)b) > h->uploaded_revision)) ++b;
        CUDA_TRY(cudaMemcpyAsync((uint32_t*)h->d_voxels + a * vol, tree.brick_pool() + a * vol, (b - a) * vol * 4,
                                 cudaMemcpyHostToDevice, h->stream));
        for (size_t k = a; k < b; ++k) handles.push_back((uint32_t)k);
        up.bricks += b - a;
        up.bytes += (b - a) * vol * 4;
        a = b;
    }
    DeviceTree& d = h->dev;
    d.node_rec = (const uint4*)h->d_node_rec;
    d.node_mip = (const uint32_t*)h->d_node_mip;
    d.mips_enabled = s.mips_enabled ? 1u : 0u;
    d.voxels = (const uint32_t*)h->d_voxels;
    d.brick_bits = (const uint32_t*)h->d_brick_bits;
    d.palette = (const uint32_t*)h->d_palette;
    d.n_nodes = (uint32_t)n_nodes;
    d.n_bricks = (uint32_t)pool;
    d.tree_size = s.tree_size;
    d.brick_dim = s.brick_dim;
    d.brick_shift = s.brick_shift;
    d.brick_dim_sq = s.brick_dim * s.brick_dim;
    d.bit_words = s.bit_words;
    d.n_colors = (uint32_t)tree.color_palette().size();
    d.inv_tree_size = 1.0f / (float)s.tree_size;
    d.inv_brick_dim = 1.0f / (float)s.brick_dim;
    // occupancy bit-bricks of the uploaded bricks, computed on the device from the resident voxels and both palettes
    if (!handles.empty()) {
        const std::vector<svx_albedo>& colors = tree.color_palette();
        const std::vector<uint32_t>& datas = tree.data_palette();
        const uint32_t color_words = (uint32_t)(colors.size() + 31) / 32, data_words = (uint32_t)(datas.size() + 31) / 32;
        std::vector<uint32_t> tables(color_words + data_words, 0u);
        for (size_t c = 0; c < colors.size(); ++c)
            if (colors[c].a != 0) tables[c >> 5] |= 1u << (c & 31);
        for (size_t k = 0; k < datas.size(); ++k)
            if (datas[k] != 0) tables[color_words + (k >> 5)] |= 1u << (k & 31);
        CUDA_TRY(grow_device_array(&h->d_data_palette, &h->data_palette_capacity, tables.size(), 4, 0, h->stream));
        CUDA_TRY(grow_device_array(&h->d_handles, &h->handle_capacity, handles.size(), 4, 0, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_data_palette, tables.data(), tables.size() * 4, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_handles, handles.data(), handles.size() * 4, cudaMemcpyHostToDevice, h->stream));
        up.bytes += (tables.size() + handles.size()) * 4;
        debug_stale("upload: before the occupancy-bit launch");
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        cudaEventRecord(e0, h->stream);
        const cudaError_t launched = launch_occupancy_bits(d, (const uint32_t*)h->d_data_palette, color_words, data_words,
                                                           (const uint32_t*)h->d_handles, (uint32_t)handles.size(),
                                                           (uint32_t*)h->d_brick_bits, h->stream);
        cudaEventRecord(e1, h->stream);
        const cudaError_t synced = cudaStreamSynchronize(h->stream);
        if (launched == cudaSuccess && synced == cudaSuccess) cudaEventElapsedTime(&up.bits_kernel_ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        CUDA_TRY(launched);
        CUDA_TRY(synced);
        h->launches += 1;
    }
    else CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stats.nodes = n_nodes;
    h->stats.bricks = s.live_bricks;
    h->stats.voxel_bytes = pool * vol * 4;
    h->stats.total_bytes = n_nodes * (sizeof(NodeHead) + 52) + s.palette.size() * 4 + pool * (vol + words) * 4;
    h->stats.tree_size = s.tree_size;
    h->stats.brick_dim = s.brick_dim;
    h->stats.depth = s.depth;
    h->stats.colours = d.n_colors;
    h->uploaded_revision = s.revision;
    h->uploaded = true;
    h->last_upload = up;
    return SVX_OK;
}

// Camera constants of reference examples/cpu_render.rs:78-103 in its f32 operation order
void make_frame_constants(const svx_view* v, FrameParams* f) {
    const svx_viewport& vp = v->viewport;
    const Vec3 origin{vp.origin[0], vp.origin[1], vp.origin[2]};
    const Vec3 dir{vp.direction[0], vp.direction[1], vp.direction[2]};
    const Vec3 up{0.0f, 1.0f, 0.0f};
    // viewport_up_direction.cross(viewport_ray.direction).normalized()  (vector.rs:186-192, :79-81)
    const Vec3 c{up.y * dir.z - up.z * dir.y, up.z * dir.x - up.x * dir.z, up.x * dir.y - up.y * dir.x};
    const float clen = std::sqrt((c.x * c.x) + (c.y * c.y) + (c.z * c.z));
    const Vec3 right = c / clen;
    const float glass_w = vp.frustum[0], glass_h = vp.frustum[1];
    const float glass_d = v->glass_mode == SVX_GLASS_AT_FRUSTUM_Z ? vp.frustum[2] : vp.fov;
    const Vec3 bl = origin + (dir * glass_d) - (up * (glass_h / 2.0f)) - (right * (glass_w / 2.0f));
    f->ox = origin.x; f->oy = origin.y; f->oz = origin.z;
    f->blx = bl.x; f->bly = bl.y; f->blz = bl.z;
    f->rx = right.x; f->ry = right.y; f->rz = right.z;
    f->ux = up.x; f->uy = up.y; f->uz = up.z;
    f->pixel_width = glass_w / (float)v->width;
    f->pixel_height = glass_h / (float)v->height;
    f->width = v->width;
    f->height = v->height;
    f->rank = v->rank;
    f->world = v->world;
    f->band_shift = 0;
    while ((1u << f->band_shift) < v->band_rows) ++f->band_shift;
    // rows owned by this shard
    uint32_t rows = 0;
    const uint32_t bands = (v->height + v->band_rows - 1) / v->band_rows;
    for (uint32_t b = v->rank; b < bands; b += v->world) rows += std::min(v->band_rows, v->height - b * v->band_rows);
    // the kernel maps local rows band by band, so a partial last band must stay addressable
    uint32_t owned_bands = 0;
    for (uint32_t b = v->rank; b < bands; b += v->world) ++owned_bands;
    f->rows_local = owned_bands * v->band_rows;
    (void)rows;
    // gather members use the rotated interleave (kernels.cuh: band_rotate): every rank walks ceil(bands / world) cycles and the
    // kernel skips the band of the last, partial cycle that falls outside the frame
    f->band_rotate = (v->gather_role != GATHER_NONE && bands < 65536u) ? 1u : 0u;
    f->world_magic = (uint32_t)((0x100000000ull + v->world - 1) / v->world);
    if (f->band_rotate) f->rows_local = ((bands + v->world - 1) / v->world) * v->band_rows;
    // Conservative screen rectangle of the root cube. The looking glass is the parallelogram
    //   origin + dir*glass_d + a*right + b*up,  a in [-w/2, w/2], b in [-h/2, h/2]   (right is perpendicular to up),
    // pixel (x, y) looks through a = -w/2 + x*pw, b = -h/2 + y*ph. A cube corner P in front of the glass plane projects
    // to the (a, b) where the line origin->P crosses that plane; rays that hit the (convex) cube pass through the convex
    // hull of the 8 projections, hence through their bounding rectangle. Done in double, padded by 2 pixels; any corner
    // not safely in front of the eye (camera inside or beside the cube) disables the cull.
    f->cull_x0 = 0; f->cull_x1 = v->width - 1; f->cull_row0 = 0; f->cull_row1 = v->height - 1;
    f->prefilter = 1u;  // no rectangle (camera inside / beside the cube): the per-ray approximate miss test earns its keep
    {
        const double S = (double)v->host->dev.tree_size;
        const double o[3] = {origin.x, origin.y, origin.z}, d[3] = {dir.x, dir.y, dir.z};
        const double rt[3] = {right.x, right.y, right.z}, u[3] = {up.x, up.y, up.z};
        const double n[3] = {rt[1] * u[2] - rt[2] * u[1], rt[2] * u[0] - rt[0] * u[2], rt[0] * u[1] - rt[1] * u[0]};
        const double dn = d[0] * n[0] + d[1] * n[1] + d[2] * n[2];
        const double rr = rt[0] * rt[0] + rt[1] * rt[1] + rt[2] * rt[2];
        double ax0 = 1e300, ax1 = -1e300, by0 = 1e300, by1 = -1e300;
        bool ok = std::isfinite(dn) && std::fabs(dn) > 1e-9 && glass_d > 0.0f && rr > 1e-12;
        for (int c = 0; c < 8 && ok; ++c) {
            const double P[3] = {(c & 1) ? S : 0.0, (c & 2) ? S : 0.0, (c & 4) ? S : 0.0};
            const double wv[3] = {P[0] - o[0], P[1] - o[1], P[2] - o[2]};
            const double wn = wv[0] * n[0] + wv[1] * n[1] + wv[2] * n[2];
            const double s = (double)glass_d * dn / wn;  // origin + s*w lies in the glass plane
            const double wl = std::sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
            if (!(wn * dn > 1e-6 * wl * std::fabs(dn)) || !std::isfinite(s) || s <= 0.0) {
                ok = false;  // corner behind / beside the eye plane
                break;
            }
            const double q[3] = {s * wv[0] - glass_d * d[0], s * wv[1] - glass_d * d[1], s * wv[2] - glass_d * d[2]};
            const double a = (q[0] * rt[0] + q[1] * rt[1] + q[2] * rt[2]) / rr;
            const double b = q[0] * u[0] + q[1] * u[1] + q[2] * u[2];
            ax0 = std::min(ax0, a); ax1 = std::max(ax1, a);
            by0 = std::min(by0, b); by1 = std::max(by1, b);
        }
        if (ok) {
            // inside a tight rectangle most rays do enter the cube: the prefilter would cost them more than it saves the others
            f->prefilter = 0u;
            const double pw = (double)f->pixel_width, ph = (double)f->pixel_height;
            const double x0 = std::floor((ax0 + glass_w / 2.0) / pw) - 2.0, x1 = std::ceil((ax1 + glass_w / 2.0) / pw) + 2.0;
            const double y0 = std::floor((by0 + glass_h / 2.0) / ph) - 2.0, y1 = std::ceil((by1 + glass_h / 2.0) / ph) + 2.0;
            const double W = v->width, H = v->height;
            if (x1 < 0 || y1 < 0 || x0 > W - 1 || y0 > H - 1) {  // nothing of the cube on screen
                f->cull_x0 = 1; f->cull_x1 = 0; f->cull_row0 = 1; f->cull_row1 = 0;
            } else {
                f->cull_x0 = (uint32_t)std::max(0.0, x0);
                f->cull_x1 = (uint32_t)std::min(W - 1, x1);
                // pixel row y lands in image row H-1-y
                f->cull_row0 = (uint32_t)(H - 1 - std::min(H - 1, y1));
                f->cull_row1 = (uint32_t)(H - 1 - std::max(0.0, y0));
            }
        }
    }
    f->compact = v->compact;
    f->viewing_distance = v->viewing_distance;
    f->shaded = v->shading ? v->d_shaded : nullptr;
    f->lx = v->light[0]; f->ly = v->light[1]; f->lz = v->light[2];
    f->go_flag = f->done_flag = f->cta_counter = nullptr;
    f->frame_seq = 0;
    f->gather_tuning = v->gather_tuning;
    f->refill_steps = v->refill_steps;
    f->refill_min_idle = v->refill_min_idle;
    f->refill_unit_tiles = v->refill_unit_tiles;
    if (v->gather_role == GATHER_PEER && !(v->gather_tuning & GATHER_TUNE_LOCAL_STORES)) {
        // the root's planes (same resolution, checked at join); 8-byte wire format: no albedo crosses NVLink
        char* base = static_cast<char*>(v->peer_block);
        f->hit_id = reinterpret_cast<uint32_t*>(base);
        f->albedo = v->gather_wire == SVX_WIRE_ID_DISTANCE ? nullptr : reinterpret_cast<uint32_t*>(base + v->peer_plane_bytes);
        f->distance = reinterpret_cast<float*>(base + 2 * v->peer_plane_bytes);
        return;
    }
    const bool alt = v->target_slot == 1;
    f->hit_id = alt ? v->alt_hit_id : v->d_hit_id;
    f->albedo = alt ? v->alt_albedo : v->d_albedo;
    f->distance = alt ? v->alt_distance : v->d_distance;
}

int32_t alloc_frame(svx_view* v) {
    cudaFree(v->frame_block);
    cudaFree(v->alt_hit_id);
    cudaFree(v->alt_albedo);
    cudaFree(v->alt_distance);
    cudaFree(v->d_shaded);
    v->frame_block = nullptr;
    v->d_hit_id = v->d_albedo = v->alt_hit_id = v->alt_albedo = v->d_shaded = nullptr;
    v->d_distance = v->alt_distance = nullptr;
    const size_t n = (size_t)v->width * v->height;
    if (v->shading) {
        CUDA_TRY(cudaMalloc((void**)&v->d_shaded, n * 4));
        CUDA_TRY(cudaMemsetAsync(v->d_shaded, 0, n * 4, v->stream));
    }
    // three planes and the gather flags in one allocation (one IPC handle exports all of it), planes 256-byte aligned
    v->plane_bytes = (n * 4 + 255) & ~(size_t)255;
    CUDA_TRY(cudaMalloc(&v->frame_block, 3 * v->plane_bytes + sizeof(GatherSync)));
    v->frame_generation += 1;
    char* base = static_cast<char*>(v->frame_block);
    v->d_hit_id = reinterpret_cast<uint32_t*>(base);
    v->d_albedo = reinterpret_cast<uint32_t*>(base + v->plane_bytes);
    v->d_distance = reinterpret_cast<float*>(base + 2 * v->plane_bytes);
    CUDA_TRY(cudaMemsetAsync(v->d_hit_id, 0xFF, v->plane_bytes, v->stream));
    CUDA_TRY(cudaMemsetAsync(v->d_albedo, 0, 2 * v->plane_bytes + sizeof(GatherSync), v->stream));
    return SVX_OK;
}

}  // namespace

namespace svx {

void octree_retain(const svx_octree* t) { t->refs.fetch_add(1); }
void octree_release(const svx_octree* t) {
    if (!t || t->refs.fetch_sub(1) != 1) return;
    svx_octree* tree = const_cast<svx_octree*>(t);
    host_release(tree->ray_host);  // holds no count on the tree: it goes with it
    delete tree->tree;
    delete tree;
}
void host_release(svx_gpu_host* h) {
    if (!h || h->refs.fetch_sub(1) != 1) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_device_tree(h);
    cudaFree(h->d_ray_lut);
    cudaFree(h->d_rays);
    cudaFree(h->d_hits);
    cudaStreamDestroy(h->stream);
    const svx_octree* tree = h->holds_octree_ref ? h->octree : nullptr;
    delete h;
    debug_stale("host_release: destroyed");
    octree_release(tree);
}

// Rejects what would make every ray of a frame NaN (the reference debug-asserts `ray.is_valid()`,
// spatial/raytracing/mod.rs:14-16, and hangs in release builds): non-finite fields, a zero direction, or a direction
// parallel to the fixed up vector (0, 1, 0) - `up x direction` is the zero vector then and its normalisation 0 / 0.
int32_t validate_viewport(const svx_viewport& vp) {
    const float all[10] = {vp.origin[0], vp.origin[1], vp.origin[2], vp.direction[0], vp.direction[1], vp.direction[2],
                           vp.frustum[0], vp.frustum[1], vp.frustum[2], vp.fov};
    for (float x : all)
        if (!std::isfinite(x)) return fail(SVX_E_INVALID_ARGUMENT, "viewport: non-finite field");
    if (vp.direction[0] == 0.0f && vp.direction[2] == 0.0f)
        return fail(SVX_E_INVALID_ARGUMENT, "viewport: the direction is zero or parallel to the up vector (0, 1, 0)");
    const float clen = std::sqrt((vp.direction[2] * vp.direction[2]) + (vp.direction[0] * vp.direction[0]));
    if (!(clen > 0.0f) || !std::isfinite(clen)) return fail(SVX_E_INVALID_ARGUMENT, "viewport: degenerate direction");
    return SVX_OK;
}

int32_t check_view_error(svx_view* v) {
    debug_stale("check_view_error");
    if (!v->h_error || *v->h_error == 0u) return SVX_OK;
    const uint32_t code = *v->h_error;
    *v->h_error = 0u;
    if (code >= 0x100u)
        return fail(SVX_E_TIMEOUT, "gather: rank " + std::to_string(code - 0x100u) + " never saw the root's go flag for this frame");
    return fail(SVX_E_TIMEOUT, "gather: peer rank " + std::to_string(code - 1u) + " did not deliver its rows in time");
}

void invalidate_block_order(svx_view* v) { v->order_valid = false; }

// Heaviest-first block order for a static-schedule launch of `f` (kernels.cuh: FrameParams::cta_order): points the launch at
// the cost array to fill and, once a frame of this shape has been recorded and sorted, at the permutation to follow.
static int32_t attach_block_order(svx_view* v, FrameParams* f, bool persistent, uint32_t* n_ctas) {
    *n_ctas = 0;
    f->cta_order = nullptr;
    f->cta_cost = nullptr;
    // OFF by default (SVX_CTA_ORDER=1: shards of a frame split four ways or more, 2: every static launch). Measured in round 2
    // (profiles/r02_schedule_probe_v*.json, r02_scaling_n8_*.json): alone on a GPU the order shortens 1/8 of sponza 4K by 6 %
    // and 1/4 of minecraft 4K by 18 %, but it costs 2-10 % on halves and whole frames (the sort is not free and neighbouring
    // blocks no longer run together), and inside the 8-GPU gather it LOSES 10 %: a sorted frame ends in a burst of short blocks
    // whose stores all cross NVLink into rank 0 at once.
    if (persistent || f->shaded || v->order_policy == 0 || (v->order_policy == 1 && v->world < 4)) return SVX_OK;
    const uint32_t n = ((f->width + 15u) / 16u) * ((f->rows_local + 7u) / 8u);  // kernels.cu: one CTA per 16x8 pixels
    if (n < 4096u && v->order_policy != 2) return SVX_OK;
    if (v->order_ctas != n) {
        cudaFree(v->d_cta_cost);
        v->d_cta_cost = v->d_cta_order = nullptr;
        v->order_ctas = 0;
        v->order_valid = false;
        CUDA_TRY(cudaMalloc((void**)&v->d_cta_cost, (size_t)2 * n * sizeof(uint32_t)));
        v->d_cta_order = v->d_cta_cost + n;
        v->order_ctas = n;
    }
    f->cta_cost = v->d_cta_cost;
    f->cta_order = v->order_valid ? v->d_cta_order : nullptr;
    *n_ctas = n;
    return SVX_OK;
}

// One frame on the view's stream. A gather peer first waits (on the device) for the root's licence to overwrite the
// shared framebuffer; a gather root follows its own rows with the kernel that waits for the peers' rows.
int32_t render_locked(svx_view* v, bool timed) {
    std::shared_lock<std::shared_mutex> tree_lock(v->host->dev_mu);
    FrameParams f;
    make_frame_constants(v, &f);
    LaunchConfig cfg = v->host->cfg;
    const bool signal_kernel = v->gather_role == GATHER_PEER && !(v->gather_tuning & (GATHER_TUNE_INKERNEL_STATIC | GATHER_TUNE_INKERNEL_PERSISTENT));
    cfg.persistent = v->persistent || (v->gather_role == GATHER_PEER && (v->gather_tuning & GATHER_TUNE_INKERNEL_PERSISTENT)) ||
                     (v->gather_role == GATHER_ROOT && (v->gather_tuning & GATHER_TUNE_PERSISTENT_ROOT));
    cfg.refill = cfg.persistent && v->refill && v->gather_role == GATHER_NONE;
    cfg.tile_counters = v->d_counters;
    // a peer's pixels cross NVLink: whole 128-byte rows from a shared-memory stage (kernels.cu: render_staged_body) once there
    // are enough senders for NVLink's delivery into rank 0 to bound the frame (measured: 8 GPUs 0.228 -> 0.168 ms with the
    // three-plane format; at 4 GPUs the stage costs 3 % and buys nothing)
    cfg.staged_stores = v->gather_role == GATHER_PEER && !cfg.persistent && v->world > 4 &&
                        !(v->gather_tuning & (GATHER_TUNE_DIRECT_STORES | GATHER_TUNE_LOCAL_STORES | GATHER_TUNE_INKERNEL_STATIC));
    f.counter_slot = v->counter_slot;
    if (cfg.persistent && !f.shaded) v->counter_slot ^= 1u;  // the shaded plane is rendered by the static schedule
    uint32_t ordered_ctas = 0;
    const int32_t attached = attach_block_order(v, &f, cfg.persistent || cfg.staged_stores, &ordered_ctas);
    if (attached != SVX_OK) return attached;
    if (v->gather_role == GATHER_PEER) {
        v->frame_seq += 1;
        GatherSync* sync = gather_sync_of(v->peer_block, v->peer_plane_bytes);
        CUDA_TRY(launch_wait_flag(&sync->go, v->frame_seq, v->gather_timeout_ns, v->h_error, 0x100u + v->rank, v->stream));
        v->launches += 1;
        f.done_flag = signal_kernel ? nullptr : &sync->done[v->rank].seq;
        f.cta_counter = v->d_cta_counter;
        f.frame_seq = v->frame_seq;
    } else if (v->gather_role == GATHER_ROOT) {
        v->frame_seq += 1;
        f.go_flag = &gather_sync_of(v->frame_block, v->plane_bytes)->go;
        f.frame_seq = v->frame_seq;
    }
    // the event pair only when the caller reads it (svx_view_render with a frame out): a timing event between two kernels of
    // a stream keeps the second from starting until the first has drained and been time-stamped - ~7 us per frame when
    // frames are queued back to back (profiles/r02_graph_probe.json)
    if (timed) CUDA_TRY(cudaEventRecord(v->ev_start, v->stream));
    CUDA_TRY(launch_render(v->host->dev, f, cfg, v->stream));
    v->launches += 1;
    if (signal_kernel) {
        CUDA_TRY(launch_signal_flag(&gather_sync_of(v->peer_block, v->peer_plane_bytes)->done[v->rank].seq, v->frame_seq, v->stream));
        v->launches += 1;
    }
    // the next frame's block order from this frame's costs. A root sorts while it waits for its peers; a peer has published
    // its rows by now, so its sort is off the frame's critical path
    if (ordered_ctas != 0 && v->gather_role != GATHER_PEER) {
        CUDA_TRY(launch_order_ctas(v->d_cta_cost, v->d_cta_order, ordered_ctas, (uint32_t)((uint64_t)ordered_ctas * v->order_head_pct / 100u), v->stream));
        v->launches += 1;
        v->order_valid = true;
    }
    if (v->gather_role == GATHER_ROOT) {
        GatherSync* sync = gather_sync_of(v->frame_block, v->plane_bytes);
        GatherComplete g{};
        g.done_flags = &sync->done[0].seq;
        g.done_stride = (uint32_t)(sizeof(sync->done[0]) / 4);
        g.world = v->world;
        g.frame_seq = v->frame_seq;
        g.fill_albedo = v->gather_wire == SVX_WIRE_ID_DISTANCE ? 1u : 0u;
        g.width = v->width;
        g.height = v->height;
        g.band_shift = f.band_shift;
        g.band_rotate = f.band_rotate;
        g.hit_id = v->d_hit_id;
        g.albedo = v->d_albedo;
        g.palette = v->host->dev.palette;
        g.n_colors = v->host->dev.n_colors;
        g.timeout_ns = v->gather_timeout_ns;
        g.error = v->h_error;
        CUDA_TRY(launch_gather_complete(g, v->host->cfg.sm_count, v->stream));
        v->launches += 1;
    }
    if (timed) CUDA_TRY(cudaEventRecord(v->ev_stop, v->stream));
    if (ordered_ctas != 0 && v->gather_role == GATHER_PEER) {
        CUDA_TRY(launch_order_ctas(v->d_cta_cost, v->d_cta_order, ordered_ctas, (uint32_t)((uint64_t)ordered_ctas * v->order_head_pct / 100u), v->stream));
        v->launches += 1;
        v->order_valid = true;
    }
    return SVX_OK;
}

// The rows this view owns, device -> host, into full-frame host planes (any may be null). A local shard (set_shard
// without a gather) copies only its own bands - strided copies whose rows are whole bands - so that several GPUs (or
// processes sharing the host buffers) assemble one frame over their own PCIe links.
int32_t copy_frame_to_host(svx_view* v, cudaStream_t stream, uint32_t* hit_id, uint32_t* albedo, float* distance) {
    if (v->gather_role == GATHER_PEER) return fail(SVX_E_INVALID_ARGUMENT, "a gather peer has no local frame to read back: read the root's");
    const void* src[3] = {v->d_hit_id, v->d_albedo, v->d_distance};
    void* dst[3] = {hit_id, albedo, distance};
    const size_t row_bytes = (size_t)v->width * 4;
    for (int p = 0; p < 3; ++p) {
        if (!dst[p]) continue;
        if (v->world == 1 || v->gather_role == GATHER_ROOT) {
            CUDA_TRY(cudaMemcpyAsync(dst[p], src[p], row_bytes * v->height, cudaMemcpyDeviceToHost, stream));
            continue;
        }
        if (v->compact) return fail(SVX_E_INVALID_ARGUMENT, "compact shards are read through svx_view_frame_pointers");
        const uint32_t band = v->band_rows, bands = (v->height + band - 1) / band;
        uint32_t full = 0;  // whole bands this shard owns; the image's last band may be partial
        for (uint32_t b = v->rank; b < bands; b += v->world)
            if ((b + 1) * band <= v->height) ++full;
        const size_t first = (size_t)v->rank * band * row_bytes, band_bytes = (size_t)band * row_bytes, pitch = band_bytes * v->world;
        if (full)
            CUDA_TRY(cudaMemcpy2DAsync((char*)dst[p] + first, pitch, (const char*)src[p] + first, pitch, band_bytes, full,
                                       cudaMemcpyDeviceToHost, stream));
        const uint32_t last = bands - 1;
        if (last % v->world == v->rank && (last + 1) * band > v->height) {
            const size_t off = (size_t)last * band_bytes;
            CUDA_TRY(cudaMemcpyAsync((char*)dst[p] + off, (const char*)src[p] + off, row_bytes * (v->height - last * band),
                                     cudaMemcpyDeviceToHost, stream));
        }
    }
    return SVX_OK;
}

// Retires pipelined frames until at most `keep` are in flight (oldest first): blocks on the slot's copy-done event and
// adds its kernel time to async_kernel_ms.
int32_t retire_locked(svx_view* v, uint32_t keep) {
    uint32_t busy = (v->slot_busy[0] ? 1u : 0u) + (v->slot_busy[1] ? 1u : 0u);
    // submission order alternates slots; the older of two busy frames sits in the slot the NEXT submission would use
    uint32_t k = (uint32_t)(v->async_frames & 1u);
    for (int n = 0; n < 2 && busy > keep; ++n, k ^= 1u) {
        if (!v->slot_busy[k]) continue;
        CUDA_TRY(cudaEventSynchronize(v->slot_copied[k]));
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, v->slot_start[k], v->slot_rendered[k]));
        v->async_kernel_ms += ms;
        v->slot_busy[k] = false;
        --busy;
    }
    return SVX_OK;
}

int32_t ensure_pipeline(svx_view* v) {
    if (!v->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CUDA_TRY(cudaEventCreate(&v->slot_start[k]));
            CUDA_TRY(cudaEventCreate(&v->slot_rendered[k]));
            CUDA_TRY(cudaEventCreateWithFlags(&v->slot_copied[k], cudaEventDisableTiming));
        }
    }
    if (!v->alt_hit_id) {
        const size_t n = (size_t)v->width * v->height;
        CUDA_TRY(cudaMalloc((void**)&v->alt_hit_id, n * 4));
        CUDA_TRY(cudaMalloc((void**)&v->alt_albedo, n * 4));
        CUDA_TRY(cudaMalloc((void**)&v->alt_distance, n * 4));
    }
    return SVX_OK;
}

// One pipelined frame: kernel into the next slot on the render stream, copies of that slot on the copy stream.
int32_t submit_async_locked(svx_view* v, uint32_t* hit_id, uint32_t* albedo, float* distance) {
    if (v->gather_role != GATHER_NONE) return fail(SVX_E_INVALID_ARGUMENT, "the pipelined read-back is not available on a gather member");
    int32_t s = ensure_pipeline(v);
    if (s != SVX_OK) return s;
    const uint32_t k = (uint32_t)(v->async_frames & 1u);
    if (v->slot_busy[k]) {  // the slot's previous frame must be on the host before the kernel overwrites it
        s = retire_locked(v, 1);
        if (s != SVX_OK) return s;
    }
    v->target_slot = k;
    std::shared_lock<std::shared_mutex> tree_lock(v->host->dev_mu);
    FrameParams f;
    make_frame_constants(v, &f);
    v->target_slot = 0;
    LaunchConfig cfg = v->host->cfg;
    cfg.persistent = v->persistent;
    cfg.refill = v->persistent && v->refill;
    cfg.tile_counters = v->d_counters;
    f.counter_slot = v->counter_slot;
    if (v->persistent && !f.shaded) v->counter_slot ^= 1u;  // the shaded plane is rendered by the static schedule
    uint32_t ordered_ctas = 0;
    const int32_t attached = attach_block_order(v, &f, cfg.persistent, &ordered_ctas);
    if (attached != SVX_OK) return attached;
    CUDA_TRY(cudaEventRecord(v->slot_start[k], v->stream));
    CUDA_TRY(launch_render(v->host->dev, f, cfg, v->stream));
    CUDA_TRY(cudaEventRecord(v->slot_rendered[k], v->stream));
    v->launches += 1;
    if (ordered_ctas != 0) {  // behind the "rendered" event: the copies of this frame do not wait for the sort
        CUDA_TRY(launch_order_ctas(v->d_cta_cost, v->d_cta_order, ordered_ctas, (uint32_t)((uint64_t)ordered_ctas * v->order_head_pct / 100u), v->stream));
        v->launches += 1;
        v->order_valid = true;
    }
    tree_lock.unlock();
    CUDA_TRY(cudaStreamWaitEvent(v->copy_stream, v->slot_rendered[k], 0));
    if (k == 0) {
        const int32_t copied = copy_frame_to_host(v, v->copy_stream, hit_id, albedo, distance);
        if (copied != SVX_OK) return copied;
    } else {  // slot 1 has its own planes: point the copy helper at them for the duration of the call
        uint32_t* const keep_hit = v->d_hit_id;
        uint32_t* const keep_alb = v->d_albedo;
        float* const keep_dist = v->d_distance;
        v->d_hit_id = v->alt_hit_id; v->d_albedo = v->alt_albedo; v->d_distance = v->alt_distance;
        const int32_t copied = copy_frame_to_host(v, v->copy_stream, hit_id, albedo, distance);
        v->d_hit_id = keep_hit; v->d_albedo = keep_alb; v->d_distance = keep_dist;
        if (copied != SVX_OK) return copied;
    }
    CUDA_TRY(cudaEventRecord(v->slot_copied[k], v->copy_stream));
    v->slot_busy[k] = true;
    v->async_frames += 1;
    return SVX_OK;
}

}  // namespace

extern "C" {

const char* svx_version(void) { return "shocovox_b200 0.1.0 (sm_100a)"; }
const char* svx_last_error_message(void) { return g_last_error.c_str(); }
int32_t svx_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ---- octree ---------------------------------------------------------------------------------------------------
int32_t svx_octree_new(uint32_t size, uint32_t brick_dim, svx_octree** out) {
    if (!out) return fail(SVX_E_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    HostOctree* t = nullptr;
    const int32_t s = HostOctree::create(size, brick_dim, &t);
    if (s != SVX_OK) return fail(s, "Octree::new rejected (size, brick_dim)");
    *out = new (std::nothrow) svx_octree();
    if (!*out) {
        delete t;
        return fail(SVX_E_OUT_OF_MEMORY, "allocation failed");
    }
    (*out)->tree = t;
    return SVX_OK;
}
void svx_octree_free(svx_octree* tree) { octree_release(tree); }
int32_t svx_octree_insert(svx_octree* t, uint32_t x, uint32_t y, uint32_t z, const svx_entry* e) {
    if (!t || !e) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    return t->tree->insert_at_lod_internal(true, x, y, z, 1, *e);
}
int32_t svx_octree_insert_at_lod(svx_octree* t, uint32_t x, uint32_t y, uint32_t z, uint32_t insert_size,
                                 const svx_entry* e) {
    if (!t || !e) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    return t->tree->insert_at_lod_internal(true, x, y, z, insert_size, *e);
}
int32_t svx_octree_update(svx_octree* t, uint32_t x, uint32_t y, uint32_t z, const svx_entry* e) {
    if (!t || !e) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    return t->tree->insert_at_lod_internal(false, x, y, z, 1, *e);
}
int32_t svx_octree_clear(svx_octree* t, uint32_t x, uint32_t y, uint32_t z) {
    if (!t) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    return t->tree->clear_at_lod(x, y, z, 1);
}
int32_t svx_octree_clear_at_lod(svx_octree* t, uint32_t x, uint32_t y, uint32_t z, uint32_t clear_size) {
    if (!t) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    return t->tree->clear_at_lod(x, y, z, clear_size);
}
int32_t svx_octree_insert_batch(svx_octree* t, const uint32_t* xyz, const uint8_t* rgba, const uint32_t* lod, uint64_t n) {
    if (!t || (n && (!xyz || !rgba))) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    for (uint64_t i = 0; i < n; ++i) {
        svx_entry e{};
        e.kind = SVX_ENTRY_VISUAL;
        e.albedo = svx_albedo{rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]};
        const uint32_t sz = (lod && lod[i] > 1) ? lod[i] : 1;
        const int32_t s = t->tree->insert_at_lod_internal(true, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], sz, e);
        if (s != SVX_OK) return fail(s, "insert failed in batch");
    }
    return SVX_OK;
}
int32_t svx_octree_get(const svx_octree* t, uint32_t x, uint32_t y, uint32_t z, svx_entry* out) {
    if (!t || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = t->tree->get(x, y, z);
    return SVX_OK;
}
int32_t svx_octree_get_sweep(const svx_octree* t, uint32_t x0, uint32_t y0, uint32_t z0, uint32_t nx, uint32_t ny,
                             uint32_t nz, svx_entry* out) {
    if (!t || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    const uint64_t size = t->tree->size();
    if ((uint64_t)x0 + nx > size || (uint64_t)y0 + ny > size || (uint64_t)z0 + nz > size)
        return fail(SVX_E_INVALID_POSITION, "the sweep box leaves the tree");
    size_t i = 0;
    for (uint64_t x = x0; x < (uint64_t)x0 + nx; ++x)
        for (uint64_t y = y0; y < (uint64_t)y0 + ny; ++y)
            for (uint64_t z = z0; z < (uint64_t)z0 + nz; ++z) out[i++] = t->tree->get((uint32_t)x, (uint32_t)y, (uint32_t)z);
    return SVX_OK;
}
uint32_t svx_octree_size(const svx_octree* t) { return t ? t->tree->size() : 0; }
uint32_t svx_octree_brick_dim(const svx_octree* t) { return t ? t->tree->brick_dim() : 0; }
int32_t svx_octree_set_auto_simplify(svx_octree* t, int32_t enabled) {
    if (!t) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    t->tree->auto_simplify = enabled != 0;
    return SVX_OK;
}
uint64_t svx_octree_structure_hash(const svx_octree* t) { return t ? t->tree->structure_hash() : 0; }
uint64_t svx_octree_node_count(const svx_octree* t) { return t ? t->tree->nodes().size() : 0; }
int32_t svx_octree_color_palette(const svx_octree* t, svx_albedo* out, uint32_t capacity, uint32_t* count) {
    if (!t || !count || (!out && capacity != 0)) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    const std::vector<svx_albedo>& pal = t->tree->color_palette();
    *count = (uint32_t)pal.size();
    if (out) std::copy_n(pal.begin(), std::min<size_t>(capacity, pal.size()), out);
    return SVX_OK;
}
int32_t svx_octree_data_palette(const svx_octree* t, uint32_t* out, uint32_t capacity, uint32_t* count) {
    if (!t || !count || (!out && capacity != 0)) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    const std::vector<uint32_t>& pal = t->tree->data_palette();
    *count = (uint32_t)pal.size();
    if (out) std::copy_n(pal.begin(), std::min<size_t>(capacity, pal.size()), out);
    return SVX_OK;
}

// ---- MIP maps: StrategyUpdater, src/octree/mipmap.rs:716-938
int32_t svx_octree_switch_albedo_mip_maps(svx_octree* t, int32_t enabled) {
    if (!t) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    t->tree->switch_albedo_mip_maps(enabled != 0);
    return SVX_OK;
}
int32_t svx_octree_mip_maps_enabled(const svx_octree* t) { return (t && t->tree->mips_enabled()) ? 1 : 0; }
int32_t svx_octree_recalculate_mips(svx_octree* t) {
    if (!t) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    t->tree->recalculate_mips();
    return SVX_OK;
}
int32_t svx_octree_mip_set_method_at(svx_octree* t, uint64_t mip_level, int32_t method, float threshold) {
    if (!t) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    if (method < SVX_MIP_BOX_FILTER || method > SVX_MIP_POSTERIZE_BD) return fail(SVX_E_INVALID_ARGUMENT, "bad MIP resampling method");
    t->tree->mip_set_method_at((size_t)mip_level, (uint32_t)method, threshold);
    return SVX_OK;
}
int32_t svx_octree_mip_get_method_at(const svx_octree* t, uint64_t mip_level, int32_t* method, float* threshold) {
    if (!t || !method) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    const MipSampler m = t->tree->mip_get_method_at((size_t)mip_level);
    *method = (int32_t)m.method;
    if (threshold) *threshold = m.thr;
    return SVX_OK;
}
int32_t svx_octree_mip_set_color_similarity_thr_at(svx_octree* t, uint64_t mip_level, float threshold) {
    if (!t) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    t->tree->mip_set_color_similarity_thr_at((size_t)mip_level, threshold);
    return SVX_OK;
}
float svx_octree_mip_get_color_similarity_at(const svx_octree* t, uint64_t mip_level) {
    return t ? t->tree->mip_get_color_similarity_at((size_t)mip_level) : 0.0f;
}
int32_t svx_octree_mip_reset(svx_octree* t) {
    if (!t) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    t->tree->mip_reset();
    return SVX_OK;
}
int32_t svx_octree_mip_sample_root(const svx_octree* t, uint32_t octant, uint32_t x, uint32_t y, uint32_t z, svx_entry* out) {
    if (!t || !out || octant > 8) return fail(SVX_E_INVALID_ARGUMENT, "bad argument");
    *out = t->tree->sample_root_mip(octant, x, y, z);
    return SVX_OK;
}
uint64_t svx_octree_mip_hash(const svx_octree* t) { return t ? t->tree->mip_hash() : 0; }

int32_t svx_octree_to_bytes(const svx_octree* t, uint8_t** bytes, uint64_t* len) {
    if (!t || !bytes || !len) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    size_t n = 0;
    uint8_t* buf = t->tree->to_bytes(&n);  // malloc'd by the encoder: handed out as it is, freed by svx_bytes_free
    if (!buf) return fail(SVX_E_OUT_OF_MEMORY, "allocation failed");
    *bytes = buf;
    *len = n;
    return SVX_OK;
}
void svx_bytes_free(uint8_t* bytes) { std::free(bytes); }
static int32_t wrap_tree(int32_t status, HostOctree* tree, svx_octree** out, const char* what) {
    if (status != SVX_OK) return fail(status, what);
    svx_octree* t = new (std::nothrow) svx_octree();
    if (!t) {
        delete tree;
        return fail(SVX_E_OUT_OF_MEMORY, "allocation failed");
    }
    t->tree = tree;
    *out = t;
    return SVX_OK;
}
int32_t svx_octree_from_bytes(const uint8_t* bytes, uint64_t len, svx_octree** out) {
    if (!bytes || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    HostOctree* tree = nullptr;
    const int32_t s = HostOctree::from_bytes(bytes, (size_t)len, &tree);
    return wrap_tree(s, tree, out, "not a bencoded Octree");
}
int32_t svx_octree_save(const svx_octree* t, const char* path) {
    if (!t || !path) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    const int32_t s = t->tree->save(path);
    return s == SVX_OK ? SVX_OK : fail(s, "cannot write the file");
}
int32_t svx_octree_load(const char* path, svx_octree** out) {
    if (!path || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    HostOctree* tree = nullptr;
    const int32_t s = HostOctree::load(path, &tree);
    return wrap_tree(s, tree, out, "cannot load an Octree from the file");
}

// ---- MagicaVoxel import: Octree::load_vox_file, src/convert/magicavoxel.rs:266-289 -----------------------------------------
int32_t svx_octree_load_vox_bytes(const uint8_t* bytes, uint64_t len, uint32_t brick_dim, svx_octree** out) {
    if (!bytes || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    HostOctree* tree = nullptr;
    std::string why;
    const int32_t s = vox_load(bytes, (size_t)len, brick_dim, &tree, &why);
    return wrap_tree(s, tree, out, why.c_str());
}
int32_t svx_octree_load_vox(const char* path, uint32_t brick_dim, svx_octree** out) {
    if (!path || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    std::vector<uint8_t> bytes;
    if (vox_read_file(path, &bytes) != SVX_OK) return fail(SVX_E_IO, "cannot read the file");
    return svx_octree_load_vox_bytes(bytes.data(), bytes.size(), brick_dim, out);
}
int32_t svx_vox_required_tree_size(const uint8_t* bytes, uint64_t len, uint32_t* tree_size) {
    if (!bytes || !tree_size) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::string why;
    const int32_t s = vox_required_tree_size(bytes, (size_t)len, tree_size, &why);
    return s == SVX_OK ? SVX_OK : fail(s, why);
}
int32_t svx_octree_insert_vox(svx_octree* t, const uint8_t* bytes, uint64_t len) {
    if (!t || !bytes) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::string why;
    const int32_t s = vox_insert_into(bytes, (size_t)len, t->tree, &why);
    return s == SVX_OK ? SVX_OK : fail(s, why);
}

// ---- gpu host -------------------------------------------------------------------------------------------------
int32_t svx_gpu_host_create(const svx_octree* tree, int32_t device, svx_gpu_host** out) {
    if (!tree || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SVX_E_CUDA, "no CUDA device available: the ray path has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(SVX_E_INVALID_ARGUMENT, "device index out of range");
    debug_stale("svx_gpu_host_create: entry");
    CUDA_TRY(cudaSetDevice(device));
    svx_gpu_host* h = new (std::nothrow) svx_gpu_host();
    if (!h) return fail(SVX_E_OUT_OF_MEMORY, "allocation failed");
    h->octree = tree;
    h->device = device;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->cfg.sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete h;
        return cuda_fail(e, "cudaStreamCreate");
    }
    std::call_once(g_selftest_once, run_selftest, h->stream);
    if (g_selftest_status != SVX_OK) {
        cudaStreamDestroy(h->stream);
        delete h;
        return fail(g_selftest_status, g_selftest_error);
    }
    {   // the ray-to-node occupancy table the traversal kernels read (validated against the device closed form above)
        std::vector<uint32_t> lut(8 * 64 * 2);
        make_ray_lut(lut.data());
        e = cudaMalloc(&h->d_ray_lut, lut.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_ray_lut, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            cudaFree(h->d_ray_lut);
            cudaStreamDestroy(h->stream);
            delete h;
            return cuda_fail(e, "ray LUT upload");
        }
        h->dev.ray_lut = (const uint2*)h->d_ray_lut;
    }
    debug_stale("svx_gpu_host_create: before upload");
    const int32_t s = upload(h);
    if (s != SVX_OK) {
        free_device_tree(h);
        cudaFree(h->d_ray_lut);
        cudaStreamDestroy(h->stream);
        delete h;
        return s;
    }
    octree_retain(tree);
    *out = h;
    return SVX_OK;
}

void svx_gpu_host_free(svx_gpu_host* h) { host_release(h); }

int32_t svx_gpu_host_reload(svx_gpu_host* h) {
    if (!h) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(h->mu);
    if (h->uploaded_revision == h->octree->tree->revision()) return SVX_OK;  // nothing was edited since the last upload
    // no view can snapshot `dev` or launch while this is held; whatever was launched before is drained below, so the
    // arrays upload() replaces or frees are no longer in use
    std::unique_lock<std::shared_mutex> tree_lock(h->dev_mu);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    return upload(h);
}

// Octree::get_by_ray / get_by_ray_at_lod on the tree handle itself (raytracing_on_cpu.rs:316-325): one ray, on the GPU.
int32_t svx_octree_get_by_ray_at_lod(svx_octree* t, const svx_ray* ray, float viewing_distance, svx_hit* out) {
    if (!t || !ray || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(t->ray_mu);
    if (!t->ray_host) {
        const int32_t created = svx_gpu_host_create(t, 0, &t->ray_host);  // SVX_E_CUDA without a device: no CPU path
        if (created != SVX_OK) return created;
        t->ray_host->holds_octree_ref = false;  // owned by the tree itself: no count, or the tree could never go
        t->refs.fetch_sub(1);
    } else {
        const int32_t reloaded = svx_gpu_host_reload(t->ray_host);
        if (reloaded != SVX_OK) return reloaded;
    }
    return svx_gpu_host_get_by_rays_at_lod(t->ray_host, ray, 1, viewing_distance, out);
}
int32_t svx_octree_get_by_ray(svx_octree* t, const svx_ray* ray, svx_hit* out) {
    return svx_octree_get_by_ray_at_lod(t, ray, 3.402823466e+38f, out);
}
// OctreeGPUView::reload, src/raytracing/bevy/mod.rs:56-60
int32_t svx_view_reload(svx_view* v) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    return svx_gpu_host_reload(v->host);
}

int32_t svx_gpu_host_last_upload(const svx_gpu_host* h, svx_upload_stats* out) {
    if (!h || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = h->last_upload;
    return SVX_OK;
}

int32_t svx_octree_render_data_nodes(const svx_octree* t, void* records, uint64_t capacity, uint64_t* n_nodes) {
    return svx_octree_render_data_nodes_with_mips(t, records, nullptr, capacity, n_nodes);
}

int32_t svx_octree_render_data_nodes_with_mips(const svx_octree* t, void* records, uint32_t* mip_slots, uint64_t capacity,
                                               uint64_t* n_nodes) {
    if (!t || !t->tree || !n_nodes) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    try {
        SerialisedNodes s;
        serialise_nodes(*t->tree, &s);
        *n_nodes = s.node_head.size();
        if (!records && !mip_slots) return SVX_OK;
        if (capacity < s.node_head.size()) return fail(SVX_E_INVALID_ARGUMENT, "record buffer too small");
        if (records) pack_node_records(s, static_cast<uint32_t*>(records));
        if (mip_slots) std::memcpy(mip_slots, s.node_mip.data(), s.node_mip.size() * 4);
        return SVX_OK;
    } catch (const std::bad_alloc&) {
        return fail(SVX_E_OUT_OF_MEMORY, "out of host memory");
    }
}

int32_t svx_octree_render_data_bricks(const svx_octree* t, uint32_t* voxels, uint32_t* bits, uint64_t capacity, uint64_t* n_bricks,
                                      uint32_t* voxels_per_brick, uint32_t* words_per_brick) {
    if (!t || !t->tree || !n_bricks || !voxels_per_brick || !words_per_brick) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    const HostOctree& tree = *t->tree;
    const size_t pool = tree.brick_pool_size(), vol = tree.brick_volume(), words = (vol + 31) / 32;
    *n_bricks = pool;
    *voxels_per_brick = (uint32_t)vol;
    *words_per_brick = (uint32_t)words;
    if (!voxels && !bits) return SVX_OK;
    if (capacity < pool) return fail(SVX_E_INVALID_ARGUMENT, "brick buffers too small");
    if (voxels && pool) std::memcpy(voxels, tree.brick_pool(), pool * vol * 4);
    if (bits) {
        // a voxel's bit is set unless pix_points_to_empty holds for it (src/octree/node.rs:405-427): (no colour index or
        // albedo.a == 0) and (no data index or data == 0) - what occupancy_bits_kernel computes on the device
        const std::vector<svx_albedo>& colors = tree.color_palette();
        const std::vector<uint32_t>& datas = tree.data_palette();
        const uint32_t* v = tree.brick_pool();
        std::memset(bits, 0, pool * words * 4);
        for (size_t b = 0; b < pool; ++b)
            for (size_t i = 0; i < vol; ++i) {
                const uint32_t value = v[b * vol + i], ci = value & 0xFFFFu, di = value >> 16;
                const bool shows = ci < colors.size() && colors[ci].a != 0, carries = di < datas.size() && datas[di] != 0;
                if (shows || carries) bits[b * words + (i >> 5)] |= 1u << (i & 31);
            }
    }
    return SVX_OK;
}

int32_t svx_render_data_ray_lut(uint32_t* lut) {
    if (!lut) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    make_ray_lut(lut);
    return SVX_OK;
}

int32_t svx_gpu_host_stats(const svx_gpu_host* h, svx_gpu_stats* out) {
    if (!h || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = h->stats;
    return SVX_OK;
}

int32_t svx_gpu_host_get_by_rays(svx_gpu_host* h, const svx_ray* rays, uint64_t n, svx_hit* hits) {
    return svx_gpu_host_get_by_rays_at_lod(h, rays, n, 3.402823466e+38f, hits);
}

int32_t svx_gpu_host_get_by_rays_at_lod(svx_gpu_host* h, const svx_ray* rays, uint64_t n, float viewing_distance,
                                        svx_hit* hits) {
    if (!h || (n && (!rays || !hits))) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    if (n == 0) return SVX_OK;
    // Ray::is_valid (spatial/raytracing/mod.rs:14-16) is a debug assertion in the reference; a NaN or zero direction would
    // make the traversal spin. Finite origins and finite non-zero directions only.
    for (uint64_t i = 0; i < n; ++i) {
        const float* o = rays[i].origin;
        const float* d = rays[i].direction;
        const bool finite = std::isfinite(o[0]) && std::isfinite(o[1]) && std::isfinite(o[2]) && std::isfinite(d[0]) &&
                            std::isfinite(d[1]) && std::isfinite(d[2]);
        if (!finite || (d[0] == 0.0f && d[1] == 0.0f && d[2] == 0.0f))
            return fail(SVX_E_INVALID_ARGUMENT, "ray " + std::to_string(i) + ": non-finite origin / direction or zero direction");
    }
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (n > h->ray_capacity) {
        cudaFree(h->d_rays);
        cudaFree(h->d_hits);
        h->d_rays = nullptr;
        h->d_hits = nullptr;
        h->ray_capacity = 0;
        CUDA_TRY(cudaMalloc((void**)&h->d_rays, n * sizeof(svx_ray)));
        CUDA_TRY(cudaMalloc((void**)&h->d_hits, n * sizeof(RayHitRecord)));
        h->ray_capacity = n;
    }
    static_assert(sizeof(svx_ray) == 24, "svx_ray is six packed floats");
    CUDA_TRY(cudaMemcpyAsync(h->d_rays, rays, n * sizeof(svx_ray), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(launch_rays(h->dev, h->d_rays, n, viewing_distance, h->d_hits, h->cfg, h->stream));
    h->launches += 1;
    std::vector<RayHitRecord> rec(n);
    CUDA_TRY(cudaMemcpyAsync(rec.data(), h->d_hits, n * sizeof(RayHitRecord), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const HostOctree& tree = *h->octree->tree;
    for (uint64_t i = 0; i < n; ++i) {
        svx_hit& o = hits[i];
        std::memset(&o, 0, sizeof(o));
        o.palette_value = rec[i].palette_value;
        o.hit = rec[i].hit;
        if (o.hit) {
            o.entry = tree.resolve(rec[i].palette_value);  // palette lookup of the returned voxel (node.rs:429-467)
            std::memcpy(o.impact_point, rec[i].impact, 12);
            std::memcpy(o.normal, rec[i].normal, 12);
            o.distance = rec[i].distance;
        } else {
            o.palette_value = NIL;
        }
    }
    return SVX_OK;
}

// ---- views ----------------------------------------------------------------------------------------------------
int32_t svx_gpu_host_create_view(svx_gpu_host* h, uint32_t, const svx_viewport* vp, uint32_t width, uint32_t height,
                                 svx_view** out) {
    if (!h || !vp || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (width == 0 || height == 0) return fail(SVX_E_INVALID_ARGUMENT, "zero resolution");
    if ((uint64_t)width * height > 0x7FFFFFFFull) return fail(SVX_E_INVALID_ARGUMENT, "resolution beyond 2^31 pixels");  // 32-bit pixel indices in the kernels
    const int32_t valid = validate_viewport(*vp);
    if (valid != SVX_OK) return valid;
    CUDA_TRY(cudaSetDevice(h->device));
    svx_view* v = new (std::nothrow) svx_view();
    if (!v) return fail(SVX_E_OUT_OF_MEMORY, "allocation failed");
    v->host = h;
    h->refs.fetch_add(1);
    v->viewport = *vp;
    v->width = width;
    v->height = height;
    cudaError_t e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&v->ev_start);
    if (e == cudaSuccess) e = cudaEventCreate(&v->ev_stop);
    if (e == cudaSuccess) e = cudaEventCreate(&v->tm_start);
    if (e == cudaSuccess) e = cudaEventCreate(&v->tm_stop);
    if (e != cudaSuccess) {
        svx_view_free(v);
        return cuda_fail(e, "view stream/event creation");
    }
    const int32_t s = alloc_frame(v);
    if (s != SVX_OK) {
        svx_view_free(v);
        return s;
    }
    e = cudaMalloc((void**)&v->d_counters, 4 * sizeof(uint32_t));  // two tile tickets + the gather's retired-CTA counter
    if (e == cudaSuccess) e = cudaMemsetAsync(v->d_counters, 0, 4 * sizeof(uint32_t), v->stream);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&v->h_error, sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) {
        svx_view_free(v);
        return cuda_fail(e, "tile counters");
    }
    *v->h_error = 0u;
    v->d_cta_counter = v->d_counters + 2;
    if (const char* o = std::getenv("SVX_CTA_ORDER")) v->order_policy = std::atoi(o);
    if (const char* o = std::getenv("SVX_CTA_ORDER_HEAD_PCT")) v->order_head_pct = (uint32_t)std::min(100, std::max(0, std::atoi(o)));
    if (const char* t = std::getenv("SVX_GATHER_TIMEOUT_MS")) v->gather_timeout_ns = (uint64_t)std::max(1L, std::atol(t)) * 1000000ull;
    const char* env = std::getenv("SVX_SCHEDULE");  // "persistent" | "static" (tuning override)
    v->persistent = env ? (std::strcmp(env, "persistent") == 0 || std::strcmp(env, "refill") == 0) : SVX_DEFAULT_PERSISTENT;
    v->refill = env && std::strcmp(env, "refill") == 0;
    if (const char* e = std::getenv("SVX_REFILL_STEPS")) v->refill_steps = std::max(1l, std::min(1l << 20, std::atol(e)));
    if (const char* e = std::getenv("SVX_REFILL_MIN_IDLE")) v->refill_min_idle = std::max(1l, std::min(32l, std::atol(e)));
    if (const char* e = std::getenv("SVX_REFILL_UNIT")) {
        const long u = std::atol(e);
        if (u == 1 || u == 2 || u == 4 || u == 8) v->refill_unit_tiles = (uint32_t)u;
    }
    *out = v;
    return SVX_OK;
}

void svx_view_free(svx_view* v) {
    if (!v) return;
    cudaSetDevice(v->host->device);
    if (v->stream) cudaStreamSynchronize(v->stream);
    svx_view_gather_close(v);
    cudaFree(v->frame_block);
    cudaFree(v->d_shaded);
    if (v->copy_stream) cudaStreamSynchronize(v->copy_stream);
    cudaFree(v->alt_hit_id);
    cudaFree(v->alt_albedo);
    cudaFree(v->alt_distance);
    for (int k = 0; k < 2; ++k) {
        if (v->slot_start[k]) cudaEventDestroy(v->slot_start[k]);
        if (v->slot_rendered[k]) cudaEventDestroy(v->slot_rendered[k]);
        if (v->slot_copied[k]) cudaEventDestroy(v->slot_copied[k]);
    }
    if (v->copy_stream) cudaStreamDestroy(v->copy_stream);
    if (v->ev_start) cudaEventDestroy(v->ev_start);
    if (v->ev_stop) cudaEventDestroy(v->ev_stop);
    if (v->tm_start) cudaEventDestroy(v->tm_start);
    if (v->tm_stop) cudaEventDestroy(v->tm_stop);
    cudaFree(v->d_flush);
    cudaFree(v->d_cta_cost);
    cudaFree(v->d_counters);
    if (v->h_error) cudaFreeHost(v->h_error);
    if (v->stream) cudaStreamDestroy(v->stream);
    svx_gpu_host* host = v->host;
    delete v;
    debug_stale("svx_view_free: exit");
    host_release(host);
}

int32_t svx_view_get_viewport(const svx_view* v, svx_viewport* out) {
    if (!v || !out) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *out = v->viewport;
    return SVX_OK;
}
int32_t svx_view_set_viewport(svx_view* v, const svx_viewport* vp) {
    if (!v || !vp) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    const int32_t valid = validate_viewport(*vp);
    if (valid != SVX_OK) return valid;
    std::lock_guard<std::mutex> lock(v->mu);
    v->viewport = *vp;
    return SVX_OK;
}
int32_t svx_view_set_glass_mode(svx_view* v, int32_t mode) {
    if (!v || (mode != SVX_GLASS_AT_FOV && mode != SVX_GLASS_AT_FRUSTUM_Z)) return fail(SVX_E_INVALID_ARGUMENT, "bad mode");
    std::lock_guard<std::mutex> lock(v->mu);
    v->glass_mode = mode;
    return SVX_OK;
}
int32_t svx_view_set_shading(svx_view* v, const float* light_normal) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    if (light_normal && v->gather_role != GATHER_NONE) return fail(SVX_E_INVALID_ARGUMENT, "the shaded plane is not gathered: leave the gather first");
    CUDA_TRY(cudaSetDevice(v->host->device));
    const int32_t drained = retire_locked(v, 0);
    if (drained != SVX_OK) return drained;
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    if (!light_normal) {
        v->shading = false;
        return SVX_OK;
    }
    std::memcpy(v->light, light_normal, 12);
    if (!v->d_shaded) {
        const size_t n = (size_t)v->width * v->height;
        CUDA_TRY(cudaMalloc((void**)&v->d_shaded, n * 4));
        CUDA_TRY(cudaMemsetAsync(v->d_shaded, 0, n * 4, v->stream));
    }
    v->shading = true;
    return SVX_OK;
}
int32_t svx_view_read_shaded(svx_view* v, uint32_t* rgba8) {
    if (!v || !rgba8) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    if (!v->shading || !v->d_shaded) return fail(SVX_E_INVALID_ARGUMENT, "shading is not enabled on this view (svx_view_set_shading)");
    CUDA_TRY(cudaSetDevice(v->host->device));
    CUDA_TRY(cudaMemcpyAsync(rgba8, v->d_shaded, (size_t)v->width * v->height * 4, cudaMemcpyDeviceToHost, v->stream));
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    return SVX_OK;
}
int32_t svx_view_shaded_pointer(const svx_view* v, void** rgba8) {
    if (!v || !rgba8) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *rgba8 = v->shading ? v->d_shaded : nullptr;
    return SVX_OK;
}
int32_t svx_view_set_viewing_distance(svx_view* v, float viewing_distance) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    v->viewing_distance = viewing_distance;
    return SVX_OK;
}
int32_t svx_view_get_viewing_distance(const svx_view* v, float* viewing_distance) {
    if (!v || !viewing_distance) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *viewing_distance = v->viewing_distance;
    return SVX_OK;
}
int32_t svx_view_set_resolution(svx_view* v, uint32_t width, uint32_t height) {
    if (!v || width == 0 || height == 0 || (uint64_t)width * height > 0x7FFFFFFFull) return fail(SVX_E_INVALID_ARGUMENT, "bad resolution");
    std::lock_guard<std::mutex> lock(v->mu);
    // peers hold mappings of (and store into) the root's frame allocation, and every member's shard assumes one resolution
    if (v->gather_role != GATHER_NONE) return fail(SVX_E_INVALID_ARGUMENT, "close the gather before changing the resolution");
    CUDA_TRY(cudaSetDevice(v->host->device));
    const int32_t drained = retire_locked(v, 0);
    if (drained != SVX_OK) return drained;
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    v->width = width;
    v->height = height;
    invalidate_block_order(v);
    return alloc_frame(v);
}
int32_t svx_view_resolution(const svx_view* v, uint32_t* width, uint32_t* height) {
    if (!v || !width || !height) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    *width = v->width;
    *height = v->height;
    return SVX_OK;
}
int32_t svx_view_set_shard(svx_view* v, uint32_t rank, uint32_t world, uint32_t rows_per_band) {
    if (!v || world == 0 || rank >= world || rows_per_band == 0 || (rows_per_band & (rows_per_band - 1)) != 0)
        return fail(SVX_E_INVALID_ARGUMENT, "bad shard (rows_per_band must be a power of two)");
    std::lock_guard<std::mutex> lock(v->mu);
    if (v->gather_role != GATHER_NONE) return fail(SVX_E_INVALID_ARGUMENT, "a gather member's shard is set by the gather");
    v->rank = rank;
    v->world = world;
    v->band_rows = rows_per_band;
    invalidate_block_order(v);
    return SVX_OK;
}

int32_t svx_view_set_schedule(svx_view* v, int32_t persistent) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    CUDA_TRY(cudaSetDevice(v->host->device));
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    CUDA_TRY(cudaMemsetAsync(v->d_counters, 0, 2 * sizeof(uint32_t), v->stream));
    v->counter_slot = 0;
    v->persistent = persistent != 0;
    v->refill = persistent == 2;
    return SVX_OK;
}

int32_t svx_view_set_compact_rows(svx_view* v, int32_t enabled) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    if (enabled && v->gather_role != GATHER_NONE) return fail(SVX_E_INVALID_ARGUMENT, "a gather member stores at image rows");
    v->compact = enabled ? 1u : 0u;
    return SVX_OK;
}

int32_t svx_view_frame_pointers(const svx_view* v, void** hit_id, void** albedo, void** distance) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    if (hit_id) *hit_id = v->d_hit_id;
    if (albedo) *albedo = v->d_albedo;
    if (distance) *distance = v->d_distance;
    return SVX_OK;
}

int32_t svx_view_render(svx_view* v, svx_frame* out) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    CUDA_TRY(cudaSetDevice(v->host->device));
    const int32_t drained = retire_locked(v, 0);
    if (drained != SVX_OK) return drained;
    const int32_t s = render_locked(v, out != nullptr);
    if (s != SVX_OK) return s;
    if (out) {
        CUDA_TRY(cudaStreamSynchronize(v->stream));
        const int32_t arrived = check_view_error(v);
        if (arrived != SVX_OK) return arrived;
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, v->ev_start, v->ev_stop));
        out->width = v->width;
        out->height = v->height;
        out->row_begin = 0;
        out->row_end = v->height;
        // a gather peer has no frame of its own: the pixels are in the root's
        const bool peer = v->gather_role == GATHER_PEER;
        out->hit_id = peer ? nullptr : v->d_hit_id;
        out->albedo = peer ? nullptr : v->d_albedo;
        out->distance = peer ? nullptr : v->d_distance;
        out->kernel_ms = ms;
    }
    return SVX_OK;
}

int32_t svx_view_render_to_host(svx_view* v, uint32_t* hit_id, uint32_t* albedo, float* distance) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    CUDA_TRY(cudaSetDevice(v->host->device));
    const int32_t drained = retire_locked(v, 0);
    if (drained != SVX_OK) return drained;
    const int32_t s = render_locked(v, false);
    if (s != SVX_OK) return s;
    if (v->gather_role != GATHER_PEER) {
        const int32_t copied = copy_frame_to_host(v, v->stream, hit_id, albedo, distance);
        if (copied != SVX_OK) return copied;
    }
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    return check_view_error(v);
}

int32_t svx_view_read_frame(svx_view* v, uint32_t* hit_id, uint32_t* albedo, float* distance) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    if (v->gather_role == GATHER_PEER) return fail(SVX_E_INVALID_ARGUMENT, "a gather peer has no local frame: read the root's");
    CUDA_TRY(cudaSetDevice(v->host->device));
    const int32_t drained = retire_locked(v, 0);
    if (drained != SVX_OK) return drained;
    const size_t bytes = (size_t)v->width * v->height * 4;
    if (hit_id) CUDA_TRY(cudaMemcpyAsync(hit_id, v->d_hit_id, bytes, cudaMemcpyDeviceToHost, v->stream));
    if (albedo) CUDA_TRY(cudaMemcpyAsync(albedo, v->d_albedo, bytes, cudaMemcpyDeviceToHost, v->stream));
    if (distance) CUDA_TRY(cudaMemcpyAsync(distance, v->d_distance, bytes, cudaMemcpyDeviceToHost, v->stream));
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    return check_view_error(v);
}

int32_t svx_view_render_to_host_async(svx_view* v, uint32_t* hit_id, uint32_t* albedo, float* distance) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    CUDA_TRY(cudaSetDevice(v->host->device));
    return submit_async_locked(v, hit_id, albedo, distance);
}

int32_t svx_view_wait_host(svx_view* v, uint32_t keep_in_flight, float* kernel_ms_total) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    CUDA_TRY(cudaSetDevice(v->host->device));
    const int32_t s = retire_locked(v, keep_in_flight);
    if (s != SVX_OK) return s;
    if (kernel_ms_total) {
        *kernel_ms_total = v->async_kernel_ms;
        v->async_kernel_ms = 0.0f;
    }
    return SVX_OK;
}

int32_t svx_view_render_batch(svx_view* v, const svx_viewport* poses, uint32_t n, uint32_t* hit_id, uint32_t* albedo,
                              float* distance, float* kernel_ms_total) {
    if (!v || (n && !poses)) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(v->mu);
    CUDA_TRY(cudaSetDevice(v->host->device));
    int32_t s = retire_locked(v, 0);
    if (s != SVX_OK) return s;
    v->async_kernel_ms = 0.0f;
    const size_t px = (size_t)v->width * v->height;
    // pose i's device->host copies overlap pose i+1's kernel (two framebuffer slots)
    for (uint32_t i = 0; i < n; ++i) {
        v->viewport = poses[i];
        s = submit_async_locked(v, hit_id ? hit_id + i * px : nullptr, albedo ? albedo + i * px : nullptr,
                                distance ? distance + i * px : nullptr);
        if (s != SVX_OK) return s;
    }
    s = retire_locked(v, 0);
    if (s != SVX_OK) return s;
    if (kernel_ms_total) *kernel_ms_total = v->async_kernel_ms;
    v->async_kernel_ms = 0.0f;
    return SVX_OK;
}

void* svx_view_cuda_stream(const svx_view* v) { return v ? (void*)v->stream : nullptr; }
int32_t svx_view_device(const svx_view* v) { return v ? v->host->device : -1; }
int32_t svx_view_synchronize(svx_view* v) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(v->host->device));
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    return check_view_error(v);
}
int32_t svx_view_timer_start(svx_view* v) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(v->host->device));
    CUDA_TRY(cudaEventRecord(v->tm_start, v->stream));
    return SVX_OK;
}
int32_t svx_view_timer_stop(svx_view* v, float* elapsed_ms) {
    if (!v || !elapsed_ms) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(v->host->device));
    CUDA_TRY(cudaEventRecord(v->tm_stop, v->stream));
    CUDA_TRY(cudaEventSynchronize(v->tm_stop));
    CUDA_TRY(cudaEventElapsedTime(elapsed_ms, v->tm_start, v->tm_stop));
    return SVX_OK;
}
int32_t svx_view_flush_l2(svx_view* v) {
    if (!v) return fail(SVX_E_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(v->host->device));
    if (!v->d_flush) {
        v->flush_bytes = (size_t)384 << 20;  // 3x the 126 MB L2
        CUDA_TRY(cudaMalloc(&v->d_flush, v->flush_bytes));
    }
    CUDA_TRY(cudaMemsetAsync(v->d_flush, (int)(v->launches & 0xFF), v->flush_bytes, v->stream));
    return SVX_OK;
}
uint64_t svx_view_launch_count(const svx_view* v) { return v ? v->launches + v->host->launches : 0; }

}  // extern "C"
