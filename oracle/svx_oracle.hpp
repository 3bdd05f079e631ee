// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// A CPU restatement (C++17, scalar f32, no FMA contraction) of the reference's
// brick-leaf sparse voxel octree and of its ray query `Octree::get_by_ray`.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this. The product (shocovox_b200/csrc) never links or calls it.
//
// Parity pinning: the reference is Rust and cannot be compiled here (no cargo/rustc),
// so this restatement is pinned against every known-answer test the reference's own
// test-suite holds for the path (tests/test_oracle_*.py):
//   src/raytracing/tests.rs:253-813 (17 literal rays), :817-911 (ring stack),
//   src/spatial/tests.rs, src/spatial/math/tests.rs, src/spatial/raytracing/tests.rs,
//   src/octree/update/tests.rs (insert / insert_at_lod / update / clear black-box tests),
//   src/octree/tests.rs:154-567 (the five mipmap KATs: MIP bricks after inserts and after recalculate_mips),
//   and the literal LUT tables in src/spatial/lut.rs:154-896 (tests/golden/luts.json).
//
// NOT pinned by any reference test (parity = the line-by-line restatement): exact impact points, the internal tree shape,
// and the level-of-detail branch of get_by_ray_at_lod with a finite viewing distance (raytracing_on_cpu.rs:368-386).
//
// All `file:line` citations are relative to /root/reference/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace svxo {

struct V3f {
    float x, y, z;
};
struct V3u {
    uint32_t x, y, z;
};

// src/spatial/mod.rs:18-39
struct Cube {
    V3f min_position;
    float size;
};

// src/spatial/raytracing/mod.rs:8-11
struct Ray {
    V3f origin;
    V3f direction;
};

// The voxels of a Parted brick: a std::vector<uint32_t> (dim^3, x fastest) that also carries a BUILD-TIME cache which never
// changes a result: which aligned 2x2x2 blocks hold more than one value, and how many do. Octree::simplify asks "is every
// block one value?" after every insert that ends in a leaf (update/mod.rs:884-980); answering it with the reference's scan
// was 77 % of the oracle's tree-build time on the 32^3-brick scenes. The cache is built by the first question and kept exact
// by the one in-place writer (the update_brick overload for Voxels, svx_oracle.cpp); assigning whole contents forgets it,
// copying a Voxels copies it along with the voxels it describes.
struct Voxels : std::vector<uint32_t> {
    static constexpr uint32_t UNKNOWN = 0xFFFFFFFFu;
    mutable std::vector<uint64_t> mixed_blocks;  // bit (bz * half + by) * half + bx
    mutable uint32_t mixed_count = UNKNOWN;

    Voxels() = default;
    Voxels(const Voxels&) = default;
    Voxels(Voxels&&) = default;
    Voxels& operator=(const Voxels&) = default;
    Voxels& operator=(Voxels&&) = default;
    Voxels& operator=(const std::vector<uint32_t>& v) {
        std::vector<uint32_t>::operator=(v);
        forget();
        return *this;
    }
    Voxels& operator=(std::vector<uint32_t>&& v) {
        std::vector<uint32_t>::operator=(std::move(v));
        forget();
        return *this;
    }
    void assign(size_t n, uint32_t v) {
        std::vector<uint32_t>::assign(n, v);
        forget();
    }
    template <class It>
    void assign(It first, It last) {
        std::vector<uint32_t>::assign(first, last);
        forget();
    }
    void resize(size_t n) {
        std::vector<uint32_t>::resize(n);
        forget();
    }
    void forget() const { mixed_count = UNKNOWN; }
};

// src/octree/types.rs:40-52
enum class BrickKind : uint8_t { Empty = 0, Parted = 1, Solid = 2 };
struct Brick {
    BrickKind kind = BrickKind::Empty;
    uint32_t solid = 0;           // valid when kind == Solid
    Voxels data;                  // valid when kind == Parted (dim^3, x fastest)
    // build-time hint only (always re-verified, never changes a result): an index that differed from data[0] the last
    // time the brick was scanned
    mutable uint32_t witness = 0;
    bool operator==(const Brick& o) const;
};

// src/octree/types.rs:56-72
enum class NodeKind : uint8_t { Nothing = 0, Internal = 1, Leaf = 2, UniformLeaf = 3 };
struct Node {
    NodeKind kind = NodeKind::Nothing;
    uint64_t occupied_bits = 0;  // Internal(u64)
    Brick bricks[8];             // Leaf([BrickData; 8])
    Brick ubrick;                // UniformLeaf(BrickData)
};

// src/octree/types.rs:76-81
enum class ChildrenKind : uint8_t { NoChildren = 0, Children = 1, OccupancyBitmap = 2 };
struct Children {
    ChildrenKind kind = ChildrenKind::NoChildren;
    uint32_t child[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t bitmap = 0;
};

// src/octree/types.rs:92-97 ; From<u32> at src/octree/detail.rs:92-105 (0xRRGGBBAA)
struct Albedo {
    uint8_t r, g, b, a;
};

// src/octree/types.rs:24-36 (T = u32)
enum class EntryKind : uint8_t { Empty = 0, Visual = 1, Informative = 2, Complex = 3 };
struct Entry {
    EntryKind kind = EntryKind::Empty;
    Albedo albedo{0, 0, 0, 0};
    uint32_t data = 0;
};

// src/octree/types.rs:9-21
enum Status : int32_t {
    OK = 0,
    E_INVALID_SIZE = 1,
    E_INVALID_BRICK_DIMENSION = 2,
    E_INVALID_STRUCTURE = 3,
    E_INVALID_POSITION = 4,
};

// src/object_pool.rs:91-241
struct NodePool {
    std::vector<Node> item;
    std::vector<uint8_t> reserved;
    size_t first_available = 0;

    size_t len() const { return item.size(); }
    bool key_is_valid(size_t key) const { return key < item.size() && reserved[key]; }
    size_t push(Node&& n);
    bool free_key(size_t key);
    void swap_items(size_t a, size_t b);

   private:
    bool is_next_available() const;
    bool check_first_available();
    size_t allocate();
};

// src/raytracing/raytracing_on_cpu.rs:20-82  NodeStack<T, SIZE = 4>: a lossy ring buffer
template <typename T, size_t SIZE = 4>
struct NodeStack {
    T data[SIZE] = {};
    size_t head_index = 0;
    uint8_t count = 0;
    bool is_empty() const { return 0 == count; }
    void push(T v) {
        head_index = (head_index + 1) % SIZE;
        count = (uint8_t)((count + 1) < (int)SIZE ? (count + 1) : (int)SIZE);
        data[head_index] = v;
    }
    bool pop(T* out) {
        if (0 == count) return false;
        count -= 1;
        if (out) *out = data[head_index];
        if (head_index == 0)
            head_index = SIZE - 1;
        else
            head_index -= 1;
        return true;
    }
    const T* last() const { return 0 == count ? nullptr : &data[head_index]; }
    T* last_mut() { return 0 == count ? nullptr : &data[head_index]; }
};

// src/octree/types.rs:104-139 MIPResamplingMethods; :149-161 MIPMapStrategy (HashMap -> ordered map: the reference never
// depends on the iteration order of these two maps except when it writes them to a file, where any order is valid)
enum class MipMethod : uint8_t { BoxFilter = 0, PointFilter = 1, PointFilterBD = 2, Posterize = 3, PosterizeBD = 4 };
struct MipSampler {
    MipMethod method = MipMethod::BoxFilter;
    float thr = 0.0f;  // Posterize / PosterizeBD parameter
};
struct MipStrategy {
    bool enabled = false;
    std::map<size_t, MipSampler> resampling_methods;
    std::map<size_t, float> resampling_color_matching_thresholds;
    MipStrategy();  // MIPMapStrategy::default(), mipmap.rs:591-604
};

struct RayStats {
    uint32_t node_iters = 0;    // iterations of the inner `while !node_stack.is_empty()` loop (raytracing_on_cpu.rs:356)
    uint32_t voxel_fetches = 0; // brick voxel reads in traverse_brick (raytracing_on_cpu.rs:218)
    uint32_t outer_iters = 0;   // iterations of `while target_octant != OOB_OCTANT` (raytracing_on_cpu.rs:352)
    uint32_t would_panic = 0;   // an index the Rust code would have bounds-panicked on
    uint32_t mip_probes = 0;    // LOD branch taken: a node's MIP brick was probed (raytracing_on_cpu.rs:377)
    uint32_t crawl_iters = 0;   // outer iterations whose node loop ran once: the root failed its occupancy test and was popped
};

struct Hit {
    bool hit = false;
    uint32_t palette_value = 0xFFFFFFFFu;  // PaletteIndexValues of the voxel returned
    Entry entry;
    V3f impact_point{0, 0, 0};
    V3f normal{0, 0, 0};
};

// src/octree/types.rs:169-207 with T = u32
class Octree {
   public:
    static Status create(uint32_t size, uint32_t brick_dim, Octree** out);

    Status insert(V3u position, const Entry& e) { return insert_at_lod_internal(true, position, 1, e); }
    Status insert_at_lod(V3u position, uint32_t size, const Entry& e) { return insert_at_lod_internal(true, position, size, e); }
    Status update(V3u position, const Entry& e) { return insert_at_lod_internal(false, position, 1, e); }
    Status clear(V3u position) { return clear_at_lod(position, 1); }   // src/octree/update/clear.rs:48-50
    Status clear_at_lod(V3u position, uint32_t clear_size);           // src/octree/update/clear.rs:55-348
    Entry get(V3u position) const;
    uint32_t get_size() const { return octree_size; }
    Hit get_by_ray(const Ray& ray, RayStats* stats = nullptr) const;  // raytracing_on_cpu.rs:316-318
    Hit get_by_ray_at_lod(const Ray& ray, float viewing_distance, RayStats* stats = nullptr) const;  // :325-565

    // ---- MIP maps: src/octree/mipmap.rs (svx_oracle_mip.cpp); StrategyUpdater methods :716-938
    MipStrategy mip_map_strategy;
    std::vector<Brick> node_mips;  // types.rs:186; always as long as the node buffer
    void switch_albedo_mip_maps(bool enabled);                       // mipmap.rs:858-872
    void recalculate_mips();                                         // mipmap.rs:798-855
    void mip_set_method_at(size_t mip_level, MipSampler method);     // mipmap.rs:657-672
    MipSampler mip_get_method_at(size_t mip_level) const;            // mipmap.rs:650-655
    void mip_set_color_similarity_thr_at(size_t mip_level, float thr);  // mipmap.rs:617-630
    float mip_get_new_color_similarity_at(size_t mip_level) const;   // mipmap.rs:610-615
    void mip_reset() { mip_map_strategy = MipStrategy(); }           // mipmap.rs:718-721
    Entry sample_root_mip(uint8_t octant, V3u position) const;       // mipmap.rs:897-937 (the reference's test hook)

    bool auto_simplify = true;

    uint32_t brick_dim = 0;
    uint32_t octree_size = 0;
    NodePool nodes;
    std::vector<Children> node_children;
    std::vector<Albedo> voxel_color_palette;
    std::vector<uint32_t> voxel_data_palette;
    // map_to_color_index_in_palette / map_to_data_index_in_palette (types.rs:194-200)
    std::unordered_map<uint32_t, size_t> color_lookup_;
    std::unordered_map<uint32_t, size_t> data_lookup_;

    // helpers exposed for tests
    uint64_t stored_occupied_bits(size_t node_key) const;
    bool pix_points_to_empty(uint32_t index) const;
    Entry pix_get_ref(uint32_t index) const;
    uint64_t structure_hash() const;  // key-order independent hash of the reachable tree
    uint64_t mip_hash() const;        // the same for the MIP strategy and the MIP bricks of the reachable nodes
    // BrickData::is_empty_throughout / is_part_empty_throughout (node.rs:107-241); public for the reference's
    // brick_tests (src/octree/tests.rs:8-151)
    bool brick_is_empty_throughout(const Brick& b, uint8_t octant) const;
    bool brick_is_part_empty_throughout(const Brick& b, uint8_t part_octant, uint8_t target_octant) const;

   private:
    Status insert_at_lod_internal(bool overwrite_if_empty, V3u position, uint32_t insert_size, const Entry& data);
    uint32_t add_to_palette(const Entry& e);
    size_t leaf_update(bool overwrite_if_empty, size_t node_key, const Cube& node_bounds, const Cube& target_bounds,
                       size_t target_child_octant, V3u position, uint32_t size, uint32_t target_content);
    void subdivide_leaf_to_nodes(size_t node_key, size_t target_octant);
    void deallocate_children_of(size_t node);
    void store_occupied_bits(size_t node_key, uint64_t bits);
    bool simplify(size_t node_key);
    Brick try_brick_from_node(size_t node_key) const;
    uint64_t brick_occupied_bits(const std::vector<uint32_t>& brick) const;
    uint64_t calculate_occupied_bits(const Brick& b) const;
    bool brick_simplify(Brick& b) const;
    bool node_is_all(const Node& n, uint32_t data) const;
    bool node_is_empty(const Node& n) const;
    bool node_empty_at(size_t node_key, uint8_t target_octant) const;
    bool should_bitmap_be_empty_at_bitmap_index(size_t node_key, size_t x, size_t y, size_t z) const;
    uint64_t hash_node(size_t key) const;
    Entry get_internal(size_t node_key, Cube bounds, V3u position) const;  // mod.rs:220-371
    void ensure_mips() { if (node_mips.size() < nodes.len()) node_mips.resize(nodes.len()); }  // insert.rs:168, detail.rs:356
    void update_mip(size_t node_key, const Cube& node_bounds, V3u position);  // mipmap.rs:296-584
    void recalculate_mip(size_t node_key, const Cube& node_bounds);           // mipmap.rs:875-892

    // ray helpers
    bool traverse_brick(const Ray& ray, V3f& p, const std::vector<uint32_t>& brick, const Cube& bounds,
                        const V3f& scale, int32_t idx_out[3], size_t& flat_out, RayStats* st) const;
    bool probe_brick(const Ray& ray, V3f& p, const Brick& brick, const Cube& bounds, const V3f& scale, Hit& out,
                     RayStats* st) const;
};

// bencode persistence, src/convert/bytecode.rs + src/octree/mod.rs:138-148 (svx_oracle_bytecode.cpp); from_bytes
// returns status 6 for bytes that are not an encoded Octree
std::string octree_to_bytes(const Octree& t);
Status octree_from_bytes(const uint8_t* data, size_t len, Octree** out);

// ---- spatial functions (exposed for known-answer tests)
uint8_t hash_region(V3f offset, float size_half);                    // src/spatial/math/mod.rs:11-19
uint8_t hash_direction(V3f direction);                               // src/spatial/math/mod.rs:22-26
size_t flat_projection(size_t x, size_t y, size_t z, size_t size);   // src/spatial/math/mod.rs:35-37
size_t position_in_bitmap_64bits(size_t x, size_t y, size_t z, size_t brick_size);  // :79-105
void set_occupancy_in_bitmap_64bits(size_t px, size_t py, size_t pz, size_t size, size_t brick_dim, bool occupied,
                                    uint64_t* bitmap);               // src/spatial/math/mod.rs:114-162
Cube child_bounds_for(const Cube& c, uint8_t octant);                // src/spatial/mod.rs:32-39
bool intersect_ray(const Cube& c, const Ray& ray, bool* has_distance, float* distance);  // src/spatial/raytracing/mod.rs:32-61
// src/spatial/raytracing/mod.rs:86-104 (not on the ray path; the reference keeps it "for debugging or new implementations")
bool plane_line_intersection(V3f plane_point, V3f plane_normal, V3f line_origin, V3f line_direction, float* distance);
V3f cross_product(V3f a, V3f b);  // V3c::cross, src/spatial/math/vector.rs:186-192 (the callers' `up x direction`)
uint8_t step_octant(uint8_t octant, V3f step);                       // src/spatial/raytracing/mod.rs:68-80
V3f cube_impact_normal(const Cube& c, V3f impact_point);             // src/spatial/raytracing/mod.rs:106-134
V3f get_dda_scale_factors(const Ray& ray);                           // src/raytracing/raytracing_on_cpu.rs:99-112
V3f dda_step_to_next_sibling(const Ray& ray, V3f& p, const Cube& bounds, const V3f& scale);  // :124-152

// LUTs regenerated from the generator logic in src/spatial/lut.rs:33-152
struct Luts {
    V3f octant_offset[8];
    uint64_t bitmap_mask_for_octant[8];
    uint8_t bitmap_index[4][4][4];
    uint32_t octant_step_result[3][3][3];
    uint64_t ray_to_node_occupancy[64][8];
};
const Luts& luts();

// Camera: caller-side ray generation of examples/cpu_render.rs:78-114
struct Camera {
    V3f origin;
    V3f direction;       // unit, looking direction
    float glass_width;   // viewport_width  (frustum.x)
    float glass_height;  // viewport_height (frustum.y)
    float glass_distance;  // viewport_fov (cpu_render.rs:92) or frustum.z (dot_cube.rs:209)
};
Ray make_pixel_ray(const Camera& cam, uint32_t w, uint32_t h, uint32_t x, uint32_t y);
// The pixel the caller loops write: examples/cpu_render.rs:119-136 (same in examples/dot_cube.rs:238-256). RGBA8 packed
// with r in the low byte, alpha 255. A hit without an albedo makes the reference panic (`albedo().unwrap()`): black here.
uint32_t shade_pixel(const Hit& hit, V3f diffuse_light_normal);

}  // namespace svxo
