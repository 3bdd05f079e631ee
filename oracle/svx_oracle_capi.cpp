// ORACLE — TEST INFRASTRUCTURE ONLY (see svx_oracle.hpp). Flat C entry points so that
// tests/, smoke() and bench.py's CPU baseline can drive the restatement through ctypes.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>

#include "svx_oracle.hpp"

using namespace svxo;

// NodeStack<i32, SIZE> KAT driver (raytracing/tests.rs:817-911): ops[i] >= 0 pushes ops[i]; -1 pops; -2 reads last;
// -3 adds 50 through last_mut. out[i] receives the popped / last value, or INT32_MIN for None. SIZE is 3 or 4.
template <size_t SIZE>
static void run_stack_script(const int32_t* ops, uint32_t n, int32_t* out) {
    NodeStack<int32_t, SIZE> stack;
    for (uint32_t i = 0; i < n; ++i) {
        out[i] = INT32_MIN;
        if (ops[i] >= 0) {
            stack.push(ops[i]);
        } else if (ops[i] == -1) {
            int32_t v;
            if (stack.pop(&v)) out[i] = v;
        } else if (ops[i] == -2) {
            if (const int32_t* l = stack.last()) out[i] = *l;
        } else {
            if (int32_t* l = stack.last_mut()) *l += 50;
        }
    }
}
extern "C" {

struct svxo_entry {
    uint32_t kind;  // 0 Empty, 1 Visual, 2 Informative, 3 Complex
    uint8_t rgba[4];
    uint32_t data;
};

struct svxo_hit {
    uint32_t hit;
    uint32_t palette_value;
    svxo_entry entry;
    float impact_point[3];
    float normal[3];
    float distance;  // (impact_point - ray.origin).length(), vector.rs:75-77
    uint32_t node_iters, voxel_fetches, outer_iters, would_panic, crawl_iters, mip_probes;
};

struct svxo_camera {
    float origin[3];
    float direction[3];
    float glass_width, glass_height, glass_distance;
};

static Entry to_entry(const svxo_entry* e) {
    Entry r;
    r.kind = (EntryKind)e->kind;
    r.albedo = Albedo{e->rgba[0], e->rgba[1], e->rgba[2], e->rgba[3]};
    r.data = e->data;
    return r;
}
static void from_entry(const Entry& e, svxo_entry* out) {
    out->kind = (uint32_t)e.kind;
    out->rgba[0] = e.albedo.r;
    out->rgba[1] = e.albedo.g;
    out->rgba[2] = e.albedo.b;
    out->rgba[3] = e.albedo.a;
    out->data = e.data;
}

int32_t svxo_octree_new(uint32_t size, uint32_t brick_dim, void** out) {
    Octree* t = nullptr;
    Status s = Octree::create(size, brick_dim, &t);
    *out = t;
    return s;
}
void svxo_octree_free(void* t) { delete (Octree*)t; }
void svxo_octree_set_auto_simplify(void* t, int32_t v) { ((Octree*)t)->auto_simplify = v != 0; }
uint32_t svxo_octree_size(void* t) { return ((Octree*)t)->get_size(); }

int32_t svxo_octree_insert(void* t, uint32_t x, uint32_t y, uint32_t z, const svxo_entry* e) {
    return ((Octree*)t)->insert(V3u{x, y, z}, to_entry(e));
}
int32_t svxo_octree_insert_at_lod(void* t, uint32_t x, uint32_t y, uint32_t z, uint32_t size, const svxo_entry* e) {
    return ((Octree*)t)->insert_at_lod(V3u{x, y, z}, size, to_entry(e));
}
int32_t svxo_octree_update(void* t, uint32_t x, uint32_t y, uint32_t z, const svxo_entry* e) {
    return ((Octree*)t)->update(V3u{x, y, z}, to_entry(e));
}
int32_t svxo_octree_clear(void* t, uint32_t x, uint32_t y, uint32_t z) { return ((Octree*)t)->clear(V3u{x, y, z}); }
int32_t svxo_octree_clear_at_lod(void* t, uint32_t x, uint32_t y, uint32_t z, uint32_t size) {
    return ((Octree*)t)->clear_at_lod(V3u{x, y, z}, size);
}
// Visual inserts in bulk: positions xyz[n][3], colours rgba[n][4]; optional per-voxel lod sizes (nullptr = 1)
int32_t svxo_octree_insert_batch(void* t, const uint32_t* xyz, const uint8_t* rgba, const uint32_t* lod, uint64_t n) {
    Octree* tree = (Octree*)t;
    for (uint64_t i = 0; i < n; ++i) {
        Entry e;
        e.kind = EntryKind::Visual;
        e.albedo = Albedo{rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]};
        const V3u p{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        const Status s = (lod && lod[i] > 1) ? tree->insert_at_lod(p, lod[i], e) : tree->insert(p, e);
        if (s != OK) return s;
    }
    return OK;
}
void svxo_octree_get(void* t, uint32_t x, uint32_t y, uint32_t z, svxo_entry* out) {
    from_entry(((Octree*)t)->get(V3u{x, y, z}), out);
}
// get() over a whole box, as palette-resolved (kind, rgba, data) triples: used for tree-equality sweeps
void svxo_octree_get_sweep(void* t, uint32_t x0, uint32_t y0, uint32_t z0, uint32_t nx, uint32_t ny, uint32_t nz,
                           svxo_entry* out) {
    Octree* tree = (Octree*)t;
    size_t i = 0;
    for (uint32_t x = x0; x < x0 + nx; ++x)
        for (uint32_t y = y0; y < y0 + ny; ++y)
            for (uint32_t z = z0; z < z0 + nz; ++z) from_entry(tree->get(V3u{x, y, z}), &out[i++]);
}
// debug aid for tests: node kinds, occupancy bits and brick kinds on the path from the root to a position
int32_t svxo_octree_describe_path(void* tp, uint32_t x, uint32_t y, uint32_t z, char* out, uint32_t cap) {
    Octree* t = (Octree*)tp;
    std::string s;
    size_t key = 0;
    float bx = 0, by = 0, bz = 0, size = (float)t->octree_size;
    for (int depth = 0; depth < 40; ++depth) {
        const Node& n = t->nodes.item[key];
        const Children& c = t->node_children[key];
        char buf[256];
        snprintf(buf, sizeof buf, "[key %zu size %g kind %d link %d ocbits %016llx]", key, size, (int)n.kind, (int)c.kind,
                 (unsigned long long)t->stored_occupied_bits(key));
        s += buf;
        const float half = size / 2;
        const int oct = ((float)x - bx >= half) + 2 * ((float)z - bz >= half) + 4 * ((float)y - by >= half);
        if (n.kind == NodeKind::Leaf) {
            s += " bricks:";
            for (int o = 0; o < 8; ++o) s += " " + std::to_string((int)n.bricks[o].kind);
            s += " target " + std::to_string(oct);
            break;
        }
        if (n.kind == NodeKind::UniformLeaf) {
            s += " ubrick " + std::to_string((int)n.ubrick.kind);
            break;
        }
        if (n.kind != NodeKind::Internal || c.kind != ChildrenKind::Children) break;
        const uint32_t ck = c.child[oct];
        s += " -> oct " + std::to_string(oct) + " child " + std::to_string(ck) + (t->nodes.key_is_valid(ck) ? "" : " (invalid)") + "\n";
        if (!t->nodes.key_is_valid(ck)) break;
        key = ck;
        bx += (oct & 1) * half; by += ((oct >> 2) & 1) * half; bz += ((oct >> 1) & 1) * half;
        size = half;
    }
    snprintf(out, cap, "%s", s.c_str());
    return (int32_t)s.size();
}
// Octree::to_bytes / from_bytes (src/octree/mod.rs:138-148). The buffer is malloc'ed: release with svxo_bytes_free.
int32_t svxo_octree_to_bytes(void* t, uint8_t** bytes, uint64_t* len) {
    const std::string s = octree_to_bytes(*(Octree*)t);
    *bytes = (uint8_t*)std::malloc(s.size() ? s.size() : 1);
    std::memcpy(*bytes, s.data(), s.size());
    *len = s.size();
    return 0;
}
void svxo_bytes_free(uint8_t* bytes) { std::free(bytes); }
int32_t svxo_octree_from_bytes(const uint8_t* bytes, uint64_t len, void** out) {
    Octree* t = nullptr;
    const Status s = octree_from_bytes(bytes, (size_t)len, &t);
    *out = t;
    return s;
}

uint64_t svxo_octree_structure_hash(void* t) { return ((Octree*)t)->structure_hash(); }
uint64_t svxo_octree_node_count(void* t) { return ((Octree*)t)->nodes.len(); }
uint64_t svxo_octree_palette_sizes(void* t, uint64_t* n_data) {
    *n_data = ((Octree*)t)->voxel_data_palette.size();
    return ((Octree*)t)->voxel_color_palette.size();
}

static void fill_hit(const Hit& h, const Ray& ray, const RayStats& st, svxo_hit* out) {
    out->hit = h.hit ? 1 : 0;
    out->palette_value = h.hit ? h.palette_value : 0xFFFFFFFFu;
    from_entry(h.entry, &out->entry);
    out->impact_point[0] = h.impact_point.x;
    out->impact_point[1] = h.impact_point.y;
    out->impact_point[2] = h.impact_point.z;
    out->normal[0] = h.normal.x;
    out->normal[1] = h.normal.y;
    out->normal[2] = h.normal.z;
    const V3f d = {h.impact_point.x - ray.origin.x, h.impact_point.y - ray.origin.y, h.impact_point.z - ray.origin.z};
    out->distance = h.hit ? std::sqrt((d.x * d.x) + (d.y * d.y) + (d.z * d.z)) : 0.0f;
    out->node_iters = st.node_iters;
    out->voxel_fetches = st.voxel_fetches;
    out->outer_iters = st.outer_iters;
    out->would_panic = st.would_panic;
    out->crawl_iters = st.crawl_iters;
    out->mip_probes = st.mip_probes;
}

// Octree::get_by_ray_at_lod (raytracing_on_cpu.rs:325); get_by_ray is viewing_distance = f32::MAX (:316-318)
void svxo_octree_get_by_ray_at_lod(void* t, const float origin[3], const float direction[3], float viewing_distance,
                                   svxo_hit* out) {
    Ray ray{{origin[0], origin[1], origin[2]}, {direction[0], direction[1], direction[2]}};
    RayStats st;
    Hit h = ((Octree*)t)->get_by_ray_at_lod(ray, viewing_distance, &st);
    fill_hit(h, ray, st, out);
}
void svxo_octree_get_by_ray(void* t, const float origin[3], const float direction[3], svxo_hit* out) {
    svxo_octree_get_by_ray_at_lod(t, origin, direction, std::numeric_limits<float>::max(), out);
}

// rays[n][6] = origin xyz, direction xyz
void svxo_octree_get_by_rays_at_lod(void* t, const float* rays, uint64_t n, float viewing_distance, svxo_hit* out) {
    for (uint64_t i = 0; i < n; ++i)
        svxo_octree_get_by_ray_at_lod(t, rays + 6 * i, rays + 6 * i + 3, viewing_distance, &out[i]);
}
void svxo_octree_get_by_rays(void* t, const float* rays, uint64_t n, svxo_hit* out) {
    svxo_octree_get_by_rays_at_lod(t, rays, n, std::numeric_limits<float>::max(), out);
}

// ---- MIP maps (src/octree/mipmap.rs): StrategyUpdater surface + the reference's sample_root_mip test hook
void svxo_octree_mip_switch(void* t, int32_t enabled) { ((Octree*)t)->switch_albedo_mip_maps(enabled != 0); }
int32_t svxo_octree_mip_enabled(void* t) { return ((Octree*)t)->mip_map_strategy.enabled ? 1 : 0; }
// method: 0 BoxFilter, 1 PointFilter, 2 PointFilterBD, 3 Posterize(thr), 4 PosterizeBD(thr)
void svxo_octree_mip_set_method_at(void* t, uint64_t level, uint32_t method, float thr) {
    ((Octree*)t)->mip_set_method_at((size_t)level, MipSampler{(MipMethod)method, thr});
}
uint32_t svxo_octree_mip_get_method_at(void* t, uint64_t level, float* thr) {
    const MipSampler m = ((Octree*)t)->mip_get_method_at((size_t)level);
    if (thr) *thr = m.thr;
    return (uint32_t)m.method;
}
void svxo_octree_mip_set_color_similarity_thr_at(void* t, uint64_t level, float thr) {
    ((Octree*)t)->mip_set_color_similarity_thr_at((size_t)level, thr);
}
float svxo_octree_mip_get_color_similarity_at(void* t, uint64_t level) {
    return ((Octree*)t)->mip_get_new_color_similarity_at((size_t)level);
}
void svxo_octree_mip_reset(void* t) { ((Octree*)t)->mip_reset(); }
void svxo_octree_mip_recalculate(void* t) { ((Octree*)t)->recalculate_mips(); }
void svxo_octree_mip_sample_root(void* t, uint32_t octant, uint32_t x, uint32_t y, uint32_t z, svxo_entry* out) {
    from_entry(((Octree*)t)->sample_root_mip((uint8_t)octant, V3u{x, y, z}), out);
}
uint64_t svxo_octree_mip_hash(void* t) { return ((Octree*)t)->mip_hash(); }

void svxo_make_pixel_ray(const svxo_camera* c, uint32_t w, uint32_t h, uint32_t x, uint32_t y, float out[6]) {
    Camera cam{{c->origin[0], c->origin[1], c->origin[2]}, {c->direction[0], c->direction[1], c->direction[2]},
               c->glass_width, c->glass_height, c->glass_distance};
    Ray r = make_pixel_ray(cam, w, h, x, y);
    out[0] = r.origin.x; out[1] = r.origin.y; out[2] = r.origin.z;
    out[3] = r.direction.x; out[4] = r.direction.y; out[5] = r.direction.z;
}

// The caller loop of examples/cpu_render.rs:104-136 over an explicit list of IMAGE rows (row = h-1-y, cpu_render.rs:106).
// Work is handed out in 64-pixel chunks from an atomic counter to `threads` host threads (0 = all hardware threads).
// Outputs (any may be null) are full [h*w] planes; only the listed rows are written. counters[6] (optional) receives
// {sum node_iters, sum voxel_fetches, sum outer_iters, rays that entered the root cube, would_panic, sum crawl_iters}.
// Returns wall-clock seconds spent in the pixel loop.
// counters[6] (7 values in the _lod variant: + MIP probes). viewing_distance as in get_by_ray_at_lod.
// shaded (optional, [h*w] RGBA8) = the pixel of the caller loop (cpu_render.rs:119-136) under light_normal[3]
double svxo_render_rows_shaded(void* t, const svxo_camera* c, uint32_t w, uint32_t h, const uint32_t* rows, uint32_t n_rows,
                               uint32_t threads, float viewing_distance, uint32_t* hit_id, uint8_t* albedo, float* distance,
                               float* normal, uint64_t* counters7, uint32_t* shaded, const float* light_normal) {
    Octree* tree = (Octree*)t;
    Camera cam{{c->origin[0], c->origin[1], c->origin[2]}, {c->direction[0], c->direction[1], c->direction[2]},
               c->glass_width, c->glass_height, c->glass_distance};
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> acc[7];
    for (auto& a : acc) a = 0;
    constexpr uint64_t CHUNK = 64;
    const uint64_t total = (uint64_t)n_rows * w;
    std::atomic<uint64_t> next{0};
    auto worker = [&]() {
        uint64_t local[7] = {0, 0, 0, 0, 0, 0, 0};
        for (;;) {
            const uint64_t begin = next.fetch_add(CHUNK);
            if (begin >= total) break;
            const uint64_t end = std::min(begin + CHUNK, total);
            for (uint64_t k = begin; k < end; ++k) {
                const uint32_t row = rows[k / w], x = (uint32_t)(k % w);
                const uint32_t y = h - 1 - row;
                const Ray ray = make_pixel_ray(cam, w, h, x, y);
                RayStats st;
                const Hit hit = tree->get_by_ray_at_lod(ray, viewing_distance, &st);
                const size_t i = (size_t)row * w + x;
                if (hit_id) hit_id[i] = hit.hit ? hit.palette_value : 0xFFFFFFFFu;
                if (albedo) {
                    const Albedo a = (hit.hit && (hit.entry.kind == EntryKind::Visual || hit.entry.kind == EntryKind::Complex))
                                         ? hit.entry.albedo
                                         : Albedo{0, 0, 0, 0};
                    albedo[4 * i] = a.r; albedo[4 * i + 1] = a.g; albedo[4 * i + 2] = a.b; albedo[4 * i + 3] = a.a;
                }
                if (distance) {
                    const V3f d = {hit.impact_point.x - ray.origin.x, hit.impact_point.y - ray.origin.y,
                                   hit.impact_point.z - ray.origin.z};
                    distance[i] = hit.hit ? std::sqrt((d.x * d.x) + (d.y * d.y) + (d.z * d.z)) : 0.0f;
                }
                if (shaded) shaded[i] = shade_pixel(hit, V3f{light_normal[0], light_normal[1], light_normal[2]});
                if (normal) {
                    normal[3 * i] = hit.hit ? hit.normal.x : 0.0f;
                    normal[3 * i + 1] = hit.hit ? hit.normal.y : 0.0f;
                    normal[3 * i + 2] = hit.hit ? hit.normal.z : 0.0f;
                }
                local[0] += st.node_iters;
                local[1] += st.voxel_fetches;
                local[2] += st.outer_iters;
                local[3] += st.outer_iters > 0 ? 1 : 0;
                local[4] += st.would_panic;
                local[5] += st.crawl_iters;
                local[6] += st.mip_probes;
            }
        }
        for (int k = 0; k < 7; ++k) acc[k] += local[k];
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (uint32_t i = 0; i < threads; ++i) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (counters7)
        for (int k = 0; k < 7; ++k) counters7[k] = acc[k];
    return std::chrono::duration<double>(t1 - t0).count();
}
double svxo_render_rows_lod(void* t, const svxo_camera* c, uint32_t w, uint32_t h, const uint32_t* rows, uint32_t n_rows,
                            uint32_t threads, float viewing_distance, uint32_t* hit_id, uint8_t* albedo, float* distance,
                            float* normal, uint64_t* counters7) {
    return svxo_render_rows_shaded(t, c, w, h, rows, n_rows, threads, viewing_distance, hit_id, albedo, distance, normal,
                                   counters7, nullptr, nullptr);
}
double svxo_render_rows(void* t, const svxo_camera* c, uint32_t w, uint32_t h, const uint32_t* rows, uint32_t n_rows,
                        uint32_t threads, uint32_t* hit_id, uint8_t* albedo, float* distance, float* normal,
                        uint64_t* counters) {
    uint64_t c7[7];
    const double s = svxo_render_rows_lod(t, c, w, h, rows, n_rows, threads, std::numeric_limits<float>::max(), hit_id,
                                          albedo, distance, normal, c7);
    if (counters)
        for (int k = 0; k < 6; ++k) counters[k] = c7[k];
    return s;
}

// Rows [row_begin, row_end) of one frame
double svxo_render(void* t, const svxo_camera* c, uint32_t w, uint32_t h, uint32_t row_begin, uint32_t row_end,
                   uint32_t threads, uint32_t* hit_id, uint8_t* albedo, float* distance, float* normal,
                   uint64_t* counters) {
    std::vector<uint32_t> rows;
    for (uint32_t r = row_begin; r < row_end; ++r) rows.push_back(r);
    return svxo_render_rows(t, c, w, h, rows.data(), (uint32_t)rows.size(), threads, hit_id, albedo, distance, normal,
                            counters);
}

uint32_t svxo_hardware_threads() { return std::max(1u, std::thread::hardware_concurrency()); }

// ---- spatial KAT entry points -------------------------------------------------------------------
uint32_t svxo_hash_region(float x, float y, float z, float half) { return hash_region(V3f{x, y, z}, half); }
uint32_t svxo_hash_direction(float x, float y, float z) { return hash_direction(V3f{x, y, z}); }
uint64_t svxo_flat_projection(uint64_t x, uint64_t y, uint64_t z, uint64_t s) { return flat_projection(x, y, z, s); }
uint64_t svxo_position_in_bitmap_64bits(uint64_t x, uint64_t y, uint64_t z, uint64_t s) {
    return position_in_bitmap_64bits(x, y, z, s);
}
uint64_t svxo_set_occupancy_in_bitmap_64bits(uint64_t x, uint64_t y, uint64_t z, uint64_t size, uint64_t dim,
                                             int32_t occupied, uint64_t bitmap) {
    set_occupancy_in_bitmap_64bits(x, y, z, size, dim, occupied != 0, &bitmap);
    return bitmap;
}
void svxo_child_bounds_for(const float min_pos[3], float size, uint32_t octant, float out[4]) {
    Cube c = child_bounds_for(Cube{{min_pos[0], min_pos[1], min_pos[2]}, size}, (uint8_t)octant);
    out[0] = c.min_position.x; out[1] = c.min_position.y; out[2] = c.min_position.z; out[3] = c.size;
}
// returns 0 = None, 1 = Some{impact_distance: None}, 2 = Some{impact_distance: Some(d)}
int32_t svxo_intersect_ray(const float min_pos[3], float size, const float origin[3], const float direction[3],
                           float* d) {
    bool has = false;
    float dist = 0;
    const bool hit = intersect_ray(Cube{{min_pos[0], min_pos[1], min_pos[2]}, size},
                                   Ray{{origin[0], origin[1], origin[2]}, {direction[0], direction[1], direction[2]}}, &has,
                                   &dist);
    *d = dist;
    return hit ? (has ? 2 : 1) : 0;
}
// returns 1 = Some(d), 0 = None
int32_t svxo_plane_line_intersection(const float plane_point[3], const float plane_normal[3], const float line_origin[3],
                                     const float line_direction[3], float* d) {
    auto v = [](const float* p) { return V3f{p[0], p[1], p[2]}; };
    return plane_line_intersection(v(plane_point), v(plane_normal), v(line_origin), v(line_direction), d) ? 1 : 0;
}
void svxo_cross(const float a[3], const float b[3], float out[3]) {
    const V3f c = cross_product(V3f{a[0], a[1], a[2]}, V3f{b[0], b[1], b[2]});
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
uint32_t svxo_step_octant(uint32_t octant, float sx, float sy, float sz) {
    return step_octant((uint8_t)octant, V3f{sx, sy, sz});
}
void svxo_cube_impact_normal(const float min_pos[3], float size, const float p[3], float out[3]) {
    V3f n = cube_impact_normal(Cube{{min_pos[0], min_pos[1], min_pos[2]}, size}, V3f{p[0], p[1], p[2]});
    out[0] = n.x; out[1] = n.y; out[2] = n.z;
}
void svxo_dda_scale_factors(const float direction[3], float out[3]) {
    V3f s = get_dda_scale_factors(Ray{{0, 0, 0}, {direction[0], direction[1], direction[2]}});
    out[0] = s.x; out[1] = s.y; out[2] = s.z;
}
// Octree::dda_step_to_next_sibling (raytracing_on_cpu.rs:124-152): advances `point` in place, writes the +-1 / 0 step
void svxo_dda_step_to_next_sibling(const float origin[3], const float direction[3], float point[3], const float min_pos[3],
                                   float size, float step_out[3]) {
    const Ray ray{{origin[0], origin[1], origin[2]}, {direction[0], direction[1], direction[2]}};
    V3f p{point[0], point[1], point[2]};
    const V3f step = dda_step_to_next_sibling(ray, p, Cube{{min_pos[0], min_pos[1], min_pos[2]}, size}, get_dda_scale_factors(ray));
    point[0] = p.x; point[1] = p.y; point[2] = p.z;
    step_out[0] = step.x; step_out[1] = step.y; step_out[2] = step.z;
}
void svxo_normalized(const float v[3], float out[3]) {
    const float len = std::sqrt((v[0] * v[0]) + (v[1] * v[1]) + (v[2] * v[2]));
    out[0] = v[0] / len; out[1] = v[1] / len; out[2] = v[2] / len;
}
// LUT dumps: mask[8], index[64] (x*16+y*4+z order of [x][y][z]), step[27] ([x][y][z]), ray2node[512] ([pos][dir])
void svxo_luts(uint64_t* mask, uint32_t* index, uint32_t* step, uint64_t* ray2node, float* offsets) {
    const Luts& l = luts();
    for (int i = 0; i < 8; ++i) mask[i] = l.bitmap_mask_for_octant[i];
    for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y)
            for (int z = 0; z < 4; ++z) index[x * 16 + y * 4 + z] = l.bitmap_index[x][y][z];
    for (int x = 0; x < 3; ++x)
        for (int y = 0; y < 3; ++y)
            for (int z = 0; z < 3; ++z) step[x * 9 + y * 3 + z] = l.octant_step_result[x][y][z];
    for (int p = 0; p < 64; ++p)
        for (int d = 0; d < 8; ++d) ray2node[p * 8 + d] = l.ray_to_node_occupancy[p][d];
    for (int o = 0; o < 8; ++o) {
        offsets[3 * o] = l.octant_offset[o].x;
        offsets[3 * o + 1] = l.octant_offset[o].y;
        offsets[3 * o + 2] = l.octant_offset[o].z;
    }
}

// ObjectPool KAT driver (src/object_pool.rs:243-285). ops[i] = {code, arg}: 0 push(arg) -> key; 1 pop(arg) -> the
// item, or INT64_MIN for None; 2 free(arg) -> 0 / 1; 3 get(arg) -> the item; 4 get_mut(arg.lo) = arg.hi -> 0;
// 5 first_available; 6 len; 7 key_is_valid(arg). Items are the Internal(u64) payload of a pooled node; `pop` is
// `free` + take (object_pool.rs:204-222).
void svxo_node_pool_script(const int64_t* ops, uint32_t n, int64_t* out) {
    NodePool pool;
    for (uint32_t i = 0; i < n; ++i) {
        const int64_t code = ops[2 * i], arg = ops[2 * i + 1];
        out[i] = INT64_MIN;
        switch (code) {
            case 0: {
                Node node;
                node.kind = NodeKind::Internal;
                node.occupied_bits = (uint64_t)arg;
                out[i] = (int64_t)pool.push(std::move(node));
                break;
            }
            case 1:
                if (pool.key_is_valid((size_t)arg)) {
                    out[i] = (int64_t)pool.item[(size_t)arg].occupied_bits;
                    pool.free_key((size_t)arg);
                    pool.item[(size_t)arg] = Node();  // std::mem::take
                }
                break;
            case 2: out[i] = pool.free_key((size_t)arg) ? 1 : 0; break;
            case 3: out[i] = (int64_t)pool.item[(size_t)arg].occupied_bits; break;
            case 4:
                pool.item[(size_t)(arg & 0xFFFFFFFF)].occupied_bits = (uint64_t)(arg >> 32);
                out[i] = 0;
                break;
            case 5: out[i] = (int64_t)pool.first_available; break;
            case 6: out[i] = (int64_t)pool.len(); break;
            case 7: out[i] = pool.key_is_valid((size_t)arg) ? 1 : 0; break;
        }
    }
}

// BrickData::is_empty_throughout (node.rs:107-178; part_octant < 0) / is_part_empty_throughout (:184-241) on a brick
// given as kind (0 Empty, 1 Parted, 2 Solid) + voxels, against the given palettes (colours as 0xRRGGBBAA).
int32_t svxo_brick_is_empty_throughout(uint32_t brick_dim, uint32_t kind, const uint32_t* voxels, uint32_t n_voxels,
                                       int32_t part_octant, uint32_t target_octant, const uint32_t* colors, uint32_t n_colors,
                                       const uint32_t* datas, uint32_t n_datas) {
    Octree* t = nullptr;
    if (Octree::create(brick_dim * 2, brick_dim, &t) != OK) return -1;
    for (uint32_t i = 0; i < n_colors; ++i)
        t->voxel_color_palette.push_back(Albedo{(uint8_t)(colors[i] >> 24), (uint8_t)(colors[i] >> 16), (uint8_t)(colors[i] >> 8), (uint8_t)colors[i]});
    t->voxel_data_palette.assign(datas, datas + n_datas);
    Brick b;
    b.kind = (BrickKind)kind;
    if (b.kind == BrickKind::Solid) b.solid = voxels[0];
    if (b.kind == BrickKind::Parted) b.data.assign(voxels, voxels + n_voxels);
    const bool empty = part_octant < 0 ? t->brick_is_empty_throughout(b, (uint8_t)target_octant)
                                       : t->brick_is_part_empty_throughout(b, (uint8_t)part_octant, (uint8_t)target_octant);
    delete t;
    return empty ? 1 : 0;
}

void svxo_node_stack_script(uint32_t size, const int32_t* ops, uint32_t n, int32_t* out) {
    if (size == 3)
        run_stack_script<3>(ops, n, out);
    else
        run_stack_script<4>(ops, n, out);
}

}  // extern "C"
