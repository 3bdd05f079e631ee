// shocovox_b200.hpp — header-only C++17 mirror of the reference crate's API for the primary-ray path, over the C ABI
// (shocovox_b200.h). Rust is not available in the build image, so this is the compiled-language host side a user of
// `shocovox_rs` switches to: same names, argument meaning and error behaviour as the crate (file:line = reference).
//
//   svx::Octree::create(size, brick_dim)            Octree::new                         src/octree/mod.rs:173
//   tree.insert / insert_at_lod / update / clear    src/octree/update/insert.rs:47-88, update/clear.rs:48-78
//   tree.get(pos)                                   Octree::get                         src/octree/mod.rs:209
//   svx::OctreeGPUHost{tree}                        OctreeGPUHost { tree }              src/raytracing/bevy/types.rs:80-87
//   host.get_by_ray(ray)                            Octree::get_by_ray                  src/raytracing/raytracing_on_cpu.rs:316
//   host.create_new_view(size, viewport, {w, h})    OctreeGPUHost::create_new_view      src/raytracing/bevy/data.rs:111
//   view.set_viewport / set_resolution / reload     OctreeSpyGlass / OctreeGPUView      src/raytracing/bevy/mod.rs:56-99
//
// Errors: the crate returns Result<_, OctreeError>; here every fallible call throws svx::OctreeError carrying the same
// variant (InvalidSize, InvalidBrickDimension, InvalidStructure, InvalidPosition) or a device error. Ray queries run on
// the GPU only; without a CUDA device OctreeGPUHost's constructor throws (there is no CPU fallback).
#pragma once
#include <array>
#include <cstdint>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "shocovox_b200.h"

namespace svx {

struct OctreeError : std::runtime_error {
    svx_status code;
    OctreeError(int32_t c, const std::string& what) : std::runtime_error(what), code(static_cast<svx_status>(c)) {}
    static const char* variant(int32_t c) {
        switch (c) {
            case SVX_E_INVALID_SIZE: return "InvalidSize";
            case SVX_E_INVALID_BRICK_DIMENSION: return "InvalidBrickDimension";
            case SVX_E_INVALID_STRUCTURE: return "InvalidStructure";
            case SVX_E_INVALID_POSITION: return "InvalidPosition";
            case SVX_E_INVALID_ARGUMENT: return "InvalidArgument";
            case SVX_E_DECODE: return "Decode";
            case SVX_E_IO: return "Io";
            case SVX_E_CUDA: return "Cuda";
            case SVX_E_OUT_OF_MEMORY: return "OutOfMemory";
            default: return "Unknown";
        }
    }
};

inline void check(int32_t status) {
    if (status != SVX_OK) throw OctreeError(status, std::string(OctreeError::variant(status)) + ": " + svx_last_error_message());
}

template <typename T>
struct V3c {  // src/spatial/math/vector.rs:9-13
    T x{}, y{}, z{};
};

struct Albedo {  // src/octree/types.rs:92-97
    uint8_t r = 0, g = 0, b = 0, a = 0;
    static Albedo from(uint32_t v) {  // Albedo::from(u32) = 0xRRGGBBAA, src/octree/detail.rs:92-105
        return Albedo{uint8_t(v >> 24), uint8_t(v >> 16), uint8_t(v >> 8), uint8_t(v)};
    }
    bool is_transparent() const { return a == 0; }
    bool operator==(const Albedo& o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
};

// OctreeEntry<u32>, src/octree/types.rs:24-36
struct OctreeEntry {
    std::optional<Albedo> albedo;
    std::optional<uint32_t> data;
    static OctreeEntry Empty() { return {}; }
    static OctreeEntry Visual(Albedo a) { return {a, std::nullopt}; }
    static OctreeEntry Informative(uint32_t d) { return {std::nullopt, d}; }
    static OctreeEntry Complex(Albedo a, uint32_t d) { return {a, d}; }
    bool is_none() const { return (!albedo || albedo->is_transparent()) && (!data || *data == 0); }  // mod.rs:102-109
    bool is_some() const { return !is_none(); }
    bool operator==(const OctreeEntry& o) const { return albedo == o.albedo && data == o.data; }

    svx_entry to_c() const {
        svx_entry e{};
        e.kind = albedo && data ? SVX_ENTRY_COMPLEX : albedo ? SVX_ENTRY_VISUAL : data ? SVX_ENTRY_INFORMATIVE : SVX_ENTRY_EMPTY;
        if (albedo) e.albedo = svx_albedo{albedo->r, albedo->g, albedo->b, albedo->a};
        if (data) e.data = *data;
        return e;
    }
    static OctreeEntry from_c(const svx_entry& e) {
        OctreeEntry r;
        if (e.kind == SVX_ENTRY_VISUAL || e.kind == SVX_ENTRY_COMPLEX) r.albedo = Albedo{e.albedo.r, e.albedo.g, e.albedo.b, e.albedo.a};
        if (e.kind == SVX_ENTRY_INFORMATIVE || e.kind == SVX_ENTRY_COMPLEX) r.data = e.data;
        return r;
    }
};

struct Ray {  // src/spatial/raytracing/mod.rs:8-11
    V3c<float> origin, direction;
};

struct Viewport {  // src/raytracing/bevy/types.rs:55-71
    V3c<float> origin, direction, frustum{4.f, 4.f, 3.f};
    float fov = 3.f;
    svx_viewport to_c() const {
        return svx_viewport{{origin.x, origin.y, origin.z}, {direction.x, direction.y, direction.z}, {frustum.x, frustum.y, frustum.z}, fov};
    }
};

struct RayHit {  // the Some((entry, impact_point, normal)) of get_by_ray
    OctreeEntry entry;
    V3c<float> impact_point, normal;
    uint32_t palette_value = 0xFFFFFFFFu;
    float distance = 0.f;
};

struct Frame {  // host copy of one rendered frame, image order (row 0 = top)
    uint32_t width = 0, height = 0;
    std::vector<uint32_t> hit_id, albedo;
    std::vector<float> distance;
};

// MIPResamplingMethods (src/octree/types.rs:106-139): the method plus, for the Posterize variants, its threshold
struct MIPResamplingMethods {
    svx_mip_method method = SVX_MIP_BOX_FILTER;
    float threshold = 0.0f;
    static MIPResamplingMethods BoxFilter() { return {SVX_MIP_BOX_FILTER, 0.0f}; }
    static MIPResamplingMethods PointFilter() { return {SVX_MIP_POINT_FILTER, 0.0f}; }
    static MIPResamplingMethods PointFilterBD() { return {SVX_MIP_POINT_FILTER_BD, 0.0f}; }
    static MIPResamplingMethods Posterize(float thr) { return {SVX_MIP_POSTERIZE, thr}; }
    static MIPResamplingMethods PosterizeBD(float thr) { return {SVX_MIP_POSTERIZE_BD, thr}; }
    bool operator==(const MIPResamplingMethods& o) const { return method == o.method && threshold == o.threshold; }
};

// StrategyUpdater<'a, T> (src/octree/types.rs:142, src/octree/mipmap.rs:716-938): chainable MIP map settings of one tree
class StrategyUpdater {
   public:
    explicit StrategyUpdater(svx_octree* tree) : h_(tree) {}
    StrategyUpdater& reset() {
        check(svx_octree_mip_reset(h_));
        return *this;
    }
    bool is_enabled() const { return svx_octree_mip_maps_enabled(h_) != 0; }
    StrategyUpdater& switch_albedo_mip_maps(bool enabled) {
        check(svx_octree_switch_albedo_mip_maps(h_, enabled ? 1 : 0));
        return *this;
    }
    StrategyUpdater& recalculate_mips() {
        check(svx_octree_recalculate_mips(h_));
        return *this;
    }
    float get_new_color_similarity_at(size_t mip_level) const { return svx_octree_mip_get_color_similarity_at(h_, mip_level); }
    StrategyUpdater& set_color_similarity_thr_at(size_t mip_level, float similarity_thr) {
        check(svx_octree_mip_set_color_similarity_thr_at(h_, mip_level, similarity_thr));
        return *this;
    }
    MIPResamplingMethods get_method_at(size_t mip_level) const {
        int32_t m = 0;
        float thr = 0.0f;
        check(svx_octree_mip_get_method_at(h_, mip_level, &m, &thr));
        return {(svx_mip_method)m, thr};
    }
    StrategyUpdater& set_method_at(size_t mip_level, MIPResamplingMethods method) {
        check(svx_octree_mip_set_method_at(h_, mip_level, method.method, method.threshold));
        return *this;
    }
    // the reference's test hook (mipmap.rs:897-937); octant 8 (OOB_OCTANT) samples the root's own MIP
    OctreeEntry sample_root_mip(uint32_t octant, V3c<uint32_t> position) const {
        svx_entry e{};
        check(svx_octree_mip_sample_root(h_, octant, position.x, position.y, position.z, &e));
        return OctreeEntry::from_c(e);
    }

   private:
    svx_octree* h_;
};

class Octree {
   public:
    static Octree create(uint32_t size, uint32_t brick_dimension) {  // Octree::new
        Octree t;
        check(svx_octree_new(size, brick_dimension, &t.h_));
        return t;
    }
    Octree(Octree&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    Octree& operator=(Octree&& o) noexcept {
        if (this != &o) {
            svx_octree_free(h_);
            h_ = std::exchange(o.h_, nullptr);
        }
        return *this;
    }
    Octree(const Octree&) = delete;
    Octree& operator=(const Octree&) = delete;
    ~Octree() { svx_octree_free(h_); }

    void insert(V3c<uint32_t> p, const OctreeEntry& e) {
        const svx_entry c = e.to_c();
        check(svx_octree_insert(h_, p.x, p.y, p.z, &c));
    }
    void insert(V3c<uint32_t> p, Albedo a) { insert(p, OctreeEntry::Visual(a)); }
    void insert_at_lod(V3c<uint32_t> p, uint32_t insert_size, const OctreeEntry& e) {
        const svx_entry c = e.to_c();
        check(svx_octree_insert_at_lod(h_, p.x, p.y, p.z, insert_size, &c));
    }
    void update(V3c<uint32_t> p, const OctreeEntry& e) {
        const svx_entry c = e.to_c();
        check(svx_octree_update(h_, p.x, p.y, p.z, &c));
    }
    void clear(V3c<uint32_t> p) { check(svx_octree_clear(h_, p.x, p.y, p.z)); }
    void clear_at_lod(V3c<uint32_t> p, uint32_t clear_size) { check(svx_octree_clear_at_lod(h_, p.x, p.y, p.z, clear_size)); }
    OctreeEntry get(V3c<uint32_t> p) const {
        svx_entry e{};
        check(svx_octree_get(h_, p.x, p.y, p.z, &e));
        return OctreeEntry::from_c(e);
    }
    uint32_t get_size() const { return svx_octree_size(h_); }
    void set_auto_simplify(bool v) { check(svx_octree_set_auto_simplify(h_, v ? 1 : 0)); }  // pub auto_simplify
    uint64_t structure_hash() const { return svx_octree_structure_hash(h_); }
    // voxel_color_palette / voxel_data_palette (types.rs:191-192): albedo of a hit = color_palette()[hit_id & 0xFFFF]
    std::vector<svx_albedo> color_palette() const {
        uint32_t n = 0;
        check(svx_octree_color_palette(h_, nullptr, 0, &n));
        std::vector<svx_albedo> colors(n);
        if (n) check(svx_octree_color_palette(h_, colors.data(), n, &n));
        return colors;
    }
    std::vector<uint32_t> data_palette() const {
        uint32_t n = 0;
        check(svx_octree_data_palette(h_, nullptr, 0, &n));
        std::vector<uint32_t> data(n);
        if (n) check(svx_octree_data_palette(h_, data.data(), n, &n));
        return data;
    }
    // Host image of the uploaded node table, 16 u32 per node (svx_octree_render_data_nodes; OctreeRenderData, bevy/types.rs:216-279)
    std::vector<std::array<uint32_t, 16>> render_data_nodes() const {
        uint64_t n = 0;
        check(svx_octree_render_data_nodes(h_, nullptr, 0, &n));
        std::vector<std::array<uint32_t, 16>> records(n);
        check(svx_octree_render_data_nodes(h_, records.data(), n, &n));
        return records;
    }
    // Octree::albedo_mip_map_resampling_strategy, src/octree/mod.rs:379
    StrategyUpdater albedo_mip_map_resampling_strategy() { return StrategyUpdater(h_); }
    // Octree::to_bytes / from_bytes / save / load (bencode, src/octree/mod.rs:138-168)
    std::vector<uint8_t> to_bytes() const {
        uint8_t* p = nullptr;
        uint64_t n = 0;
        check(svx_octree_to_bytes(h_, &p, &n));
        std::vector<uint8_t> out(p, p + n);
        svx_bytes_free(p);
        return out;
    }
    static Octree from_bytes(const std::vector<uint8_t>& bytes) {
        Octree t;
        check(svx_octree_from_bytes(bytes.data(), bytes.size(), &t.h_));
        return t;
    }
    void save(const std::string& path) const { check(svx_octree_save(h_, path.c_str())); }
    static Octree load(const std::string& path) {
        Octree t;
        check(svx_octree_load(path.c_str(), &t.h_));
        return t;
    }
    // Octree::load_vox_file(filename, brick_dimension), src/convert/magicavoxel.rs:266-289
    static Octree load_vox_file(const std::string& filename, uint32_t brick_dimension) {
        Octree t;
        check(svx_octree_load_vox(filename.c_str(), brick_dimension, &t.h_));
        return t;
    }
    // MIPMapStrategy::...::load_vox_file (magicavoxel.rs:207-250): `configure` sets the MIP strategy of the still EMPTY tree
    // (e.g. `[](StrategyUpdater s) { s.switch_albedo_mip_maps(true); }`), then the voxels go in and refresh the MIPs as they do
    template <typename Configure>
    static Octree load_vox_file(const std::vector<uint8_t>& vox_bytes, uint32_t brick_dimension, Configure&& configure) {
        uint32_t tree_size = 0;
        check(svx_vox_required_tree_size(vox_bytes.data(), vox_bytes.size(), &tree_size));
        Octree t = create(tree_size, brick_dimension);
        configure(t.albedo_mip_map_resampling_strategy());
        check(svx_octree_insert_vox(t.h_, vox_bytes.data(), vox_bytes.size()));
        return t;
    }
    svx_octree* handle() const { return h_; }

   private:
    Octree() = default;
    svx_octree* h_ = nullptr;
};

class OctreeGPUView;

class OctreeGPUHost {  // OctreeGPUHost { tree }: uploads the whole tree to `device`
   public:
    explicit OctreeGPUHost(const Octree& tree, int device = 0) { check(svx_gpu_host_create(tree.handle(), device, &h_)); }
    OctreeGPUHost(OctreeGPUHost&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    OctreeGPUHost(const OctreeGPUHost&) = delete;
    OctreeGPUHost& operator=(const OctreeGPUHost&) = delete;
    ~OctreeGPUHost() { svx_gpu_host_free(h_); }

    void reload() { check(svx_gpu_host_reload(h_)); }  // incremental: node tables + the bricks written since the last upload
    svx_upload_stats last_upload() const {
        svx_upload_stats s{};
        check(svx_gpu_host_last_upload(h_, &s));
        return s;
    }
    svx_gpu_stats stats() const {
        svx_gpu_stats s{};
        check(svx_gpu_host_stats(h_, &s));
        return s;
    }
    // Octree::get_by_ray(&Ray) -> Option<(OctreeEntry, V3c<f32>, V3c<f32>)>, on the GPU
    std::optional<RayHit> get_by_ray(const Ray& ray) { return get_by_ray_at_lod(ray, std::numeric_limits<float>::max()); }
    // Octree::get_by_ray_at_lod(&Ray, viewing_distance), src/raytracing/raytracing_on_cpu.rs:325
    std::optional<RayHit> get_by_ray_at_lod(const Ray& ray, float viewing_distance) {
        std::vector<std::optional<RayHit>> r = get_by_rays({ray}, viewing_distance);
        return r[0];
    }
    std::vector<std::optional<RayHit>> get_by_rays(const std::vector<Ray>& rays,
                                                   float viewing_distance = std::numeric_limits<float>::max()) {
        std::vector<svx_ray> in(rays.size());
        for (size_t i = 0; i < rays.size(); ++i)
            in[i] = svx_ray{{rays[i].origin.x, rays[i].origin.y, rays[i].origin.z},
                            {rays[i].direction.x, rays[i].direction.y, rays[i].direction.z}};
        std::vector<svx_hit> out(rays.size());
        check(svx_gpu_host_get_by_rays_at_lod(h_, in.data(), in.size(), viewing_distance, out.data()));
        std::vector<std::optional<RayHit>> res(rays.size());
        for (size_t i = 0; i < rays.size(); ++i) {
            if (!out[i].hit) continue;
            RayHit h;
            h.entry = OctreeEntry::from_c(out[i].entry);
            h.impact_point = {out[i].impact_point[0], out[i].impact_point[1], out[i].impact_point[2]};
            h.normal = {out[i].normal[0], out[i].normal[1], out[i].normal[2]};
            h.palette_value = out[i].palette_value;
            h.distance = out[i].distance;
            res[i] = h;
        }
        return res;
    }
    inline OctreeGPUView create_new_view(uint32_t size, const Viewport& viewport, std::array<uint32_t, 2> resolution);
    svx_gpu_host* handle() const { return h_; }

   private:
    svx_gpu_host* h_ = nullptr;
};

class OctreeGPUView {
   public:
    OctreeGPUView(OctreeGPUHost& host, uint32_t size, const Viewport& vp, std::array<uint32_t, 2> res) : host_(&host) {
        const svx_viewport c = vp.to_c();
        check(svx_gpu_host_create_view(host.handle(), size, &c, res[0], res[1], &h_));
    }
    OctreeGPUView(OctreeGPUView&& o) noexcept : host_(o.host_), h_(std::exchange(o.h_, nullptr)) {}
    OctreeGPUView(const OctreeGPUView&) = delete;
    OctreeGPUView& operator=(const OctreeGPUView&) = delete;
    ~OctreeGPUView() { svx_view_free(h_); }

    void reload() { host_->reload(); }                      // OctreeGPUView::reload, bevy/mod.rs:56-60
    void set_viewport(const Viewport& vp) {                 // spyglass.viewport_mut()
        const svx_viewport c = vp.to_c();
        check(svx_view_set_viewport(h_, &c));
    }
    Viewport viewport() const {                             // spyglass.viewport()
        svx_viewport c{};
        check(svx_view_get_viewport(h_, &c));
        return Viewport{{c.origin[0], c.origin[1], c.origin[2]}, {c.direction[0], c.direction[1], c.direction[2]},
                        {c.frustum[0], c.frustum[1], c.frustum[2]}, c.fov};
    }
    void set_glass_mode(svx_glass_mode m) { check(svx_view_set_glass_mode(h_, m)); }
    // viewing distance of every pixel's get_by_ray_at_lod (default f32::MAX = get_by_ray; the reference's GPU path
    // feeds viewport.frustum.z); only matters while the tree's MIP maps are enabled
    // the caller loops' shaded pixel (examples/cpu_render.rs:119-136) as a fourth plane; see svx_view_set_shading
    void set_shading(V3c<float> diffuse_light_normal) {
        const float l[3] = {diffuse_light_normal.x, diffuse_light_normal.y, diffuse_light_normal.z};
        check(svx_view_set_shading(h_, l));
    }
    void disable_shading() { check(svx_view_set_shading(h_, nullptr)); }
    std::vector<uint32_t> read_shaded() {  // RGBA8 of the last rendered frame, r in the low byte
        const auto r = resolution();
        std::vector<uint32_t> out(size_t(r[0]) * r[1]);
        check(svx_view_read_shaded(h_, out.data()));
        return out;
    }
    void set_viewing_distance(float d) { check(svx_view_set_viewing_distance(h_, d)); }
    float viewing_distance() const {
        float d = 0.0f;
        check(svx_view_get_viewing_distance(h_, &d));
        return d;
    }
    void set_resolution(std::array<uint32_t, 2> r) { check(svx_view_set_resolution(h_, r[0], r[1])); }  // bevy/mod.rs:62-82
    std::array<uint32_t, 2> resolution() const {
        uint32_t w = 0, h = 0;
        check(svx_view_resolution(h_, &w, &h));
        return {w, h};
    }
    void set_shard(uint32_t rank, uint32_t world, uint32_t rows_per_band = 8) { check(svx_view_set_shard(h_, rank, world, rows_per_band)); }
    // one frame on the device; the returned pointers stay owned by the view
    svx_frame render() {
        svx_frame f{};
        check(svx_view_render(h_, &f));
        return f;
    }
    // one frame, copied to the host
    Frame render_to_host() {
        Frame f;
        const auto r = resolution();
        f.width = r[0];
        f.height = r[1];
        f.hit_id.resize(size_t(r[0]) * r[1]);
        f.albedo.resize(f.hit_id.size());
        f.distance.resize(f.hit_id.size());
        check(svx_view_render_to_host(h_, f.hit_id.data(), f.albedo.data(), f.distance.data()));
        return f;
    }
    // pipelined frame into caller-owned (pinned) host planes; complete after wait_host() retired it
    void render_to_host_async(uint32_t* hit_id, uint32_t* albedo, float* distance) {
        check(svx_view_render_to_host_async(h_, hit_id, albedo, distance));
    }
    float wait_host(uint32_t keep_in_flight = 0) {
        float ms = 0.0f;
        check(svx_view_wait_host(h_, keep_in_flight, &ms));
        return ms;
    }
    // the framebuffer as it stands (on the root of a gather: the assembled frame of the last render)
    Frame read_frame() {
        Frame f;
        const auto r = resolution();
        f.width = r[0];
        f.height = r[1];
        f.hit_id.resize(size_t(r[0]) * r[1]);
        f.albedo.resize(f.hit_id.size());
        f.distance.resize(f.hit_id.size());
        check(svx_view_read_frame(h_, f.hit_id.data(), f.albedo.data(), f.distance.data()));
        return f;
    }
    // one frame over several GPUs, one process per GPU (the reference renders on one device; see svx_view_gather_* in the C
    // header): the root's framebuffer assembles the frame, a peer's kernel stores into it
    svx_gather_handle gather_open(uint32_t world, uint32_t rows_per_band = 8, svx_wire_format wire = SVX_WIRE_THREE_PLANES) {
        svx_gather_handle h{};
        check(svx_view_gather_open(h_, world, rows_per_band, wire, &h));
        return h;
    }
    void gather_join(uint32_t rank, const svx_gather_handle& handle) { check(svx_view_gather_join(h_, rank, &handle)); }
    void gather_join_local(uint32_t rank, OctreeGPUView& root) { check(svx_view_gather_join_local(h_, rank, root.h_)); }
    void gather_close() { check(svx_view_gather_close(h_)); }
    svx_view* handle() const { return h_; }

   private:
    OctreeGPUHost* host_;
    svx_view* h_ = nullptr;
};

// One process driving several GPUs (svx_multi_*): a replica of the tree and a view per device, ONE frame per render call
class MultiGPU {
   public:
    MultiGPU(const Octree& tree, const std::vector<int32_t>& devices, const Viewport& vp, std::array<uint32_t, 2> res,
             uint32_t rows_per_band = 8, svx_wire_format wire = SVX_WIRE_THREE_PLANES)
        : res_(res) {
        const svx_viewport c = vp.to_c();
        check(svx_multi_create(tree.handle(), devices.data(), (uint32_t)devices.size(), &c, res[0], res[1], rows_per_band, wire, &h_));
    }
    MultiGPU(const MultiGPU&) = delete;
    MultiGPU& operator=(const MultiGPU&) = delete;
    ~MultiGPU() { svx_multi_free(h_); }
    void set_viewport(const Viewport& vp) {
        const svx_viewport c = vp.to_c();
        check(svx_multi_set_viewport(h_, &c));
    }
    void set_glass_mode(svx_glass_mode m) { check(svx_multi_set_glass_mode(h_, m)); }
    void set_viewing_distance(float d) { check(svx_multi_set_viewing_distance(h_, d)); }
    void reload() { check(svx_multi_reload(h_)); }
    svx_frame render() {  // gathered in devices[0]'s framebuffer by the viewport kernels themselves
        svx_frame f{};
        check(svx_multi_render(h_, &f));
        return f;
    }
    Frame render_to_host() {  // every GPU copies the rows it rendered over its own PCIe link
        Frame f;
        f.width = res_[0];
        f.height = res_[1];
        f.hit_id.resize(size_t(res_[0]) * res_[1]);
        f.albedo.resize(f.hit_id.size());
        f.distance.resize(f.hit_id.size());
        check(svx_multi_render_to_host(h_, f.hit_id.data(), f.albedo.data(), f.distance.data()));
        return f;
    }

   private:
    std::array<uint32_t, 2> res_;
    svx_multi* h_ = nullptr;
};

inline OctreeGPUView OctreeGPUHost::create_new_view(uint32_t size, const Viewport& viewport, std::array<uint32_t, 2> resolution) {
    return OctreeGPUView(*this, size, viewport, resolution);
}

}  // namespace svx
