// Host-side brick-leaf sparse voxel octree of the product: the construction API the reference keeps on the CPU
// (Octree::new / insert / insert_at_lod / update / get, reference src/octree/**), stored the way the GPU wants it.
//
// Layout: fixed-size node records in one arena plus ONE pooled voxel array for all parted bricks (a brick is a
// handle into the pool, not an owned Vec as in the reference's BrickData::Parted, src/octree/types.rs:40-52).
// A node record already carries what the device record needs (occupancy bits, brick kinds, 8 slots), so the
// render-data upload (gpu_tree.cpp) is a breadth-first compaction, not a re-encoding.
//
// The tree *shape* decides which f32 path a ray takes, so every mutation follows the reference's
// rules exactly; tests/test_host_octree.py checks shape equality against the oracle with a structural hash.
// All `file:line` citations are relative to the reference checkout.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/shocovox_b200.h"

namespace svx {

constexpr uint32_t NIL = 0xFFFFFFFFu;  // empty_marker::<u32>(), src/object_pool.rs:14-16

enum NodeKind : uint8_t { NK_NOTHING = 0, NK_INTERNAL = 1, NK_LEAF = 2, NK_UNIFORM = 3 };  // NodeContent, types.rs:56-72
enum LinkKind : uint8_t { LK_NONE = 0, LK_CHILDREN = 1, LK_BITMAP = 2 };                  // NodeChildren, types.rs:76-81
enum BrickKind : uint8_t { BK_EMPTY = 0, BK_PARTED = 1, BK_SOLID = 2 };                   // BrickData, types.rs:40-52

// A brick reference: Solid carries the palette value itself, Parted a handle into the voxel pool.
struct BrickRef {
    uint8_t kind = BK_EMPTY;
    uint32_t value = NIL;
};

struct NodeRec {
    uint64_t ocbits = 0;     // NodeContent::Internal(u64)
    uint64_t leaf_bits = 0;  // NodeChildren::OccupancyBitmap(u64)
    uint32_t child[8] = {NIL, NIL, NIL, NIL, NIL, NIL, NIL, NIL};  // NodeChildren::Children
    BrickRef brick[8];       // Leaf: one per octant ; UniformLeaf: brick[0]
    uint8_t kind = NK_NOTHING;
    uint8_t link = LK_NONE;
    uint8_t reserved = 0;    // ObjectPool's `reserved` flag, src/object_pool.rs:9-12
    // node_mips[key] (types.rs:186): a brick of the same pool holding the node's simplified view. It belongs to the
    // KEY, not to the node content: the reference never resets it when a key is freed and reused, so neither
    // clear_content() nor pool_free() touch it
    BrickRef mip;
};

// MIPResamplingMethods (types.rs:106-139); the numbering is the C ABI's SVX_MIP_*
enum MipMethod : uint32_t { MIP_BOX = 0, MIP_POINT = 1, MIP_POINT_BD = 2, MIP_POSTERIZE = 3, MIP_POSTERIZE_BD = 4 };
struct MipSampler {
    uint32_t method = MIP_BOX;
    float thr = 0.0f;
};

struct BoundsF {  // Cube, src/spatial/mod.rs:18-21
    float x, y, z, size;
};

class HostOctree {
   public:
    static int32_t create(uint32_t size, uint32_t brick_dim, HostOctree** out);

    int32_t insert_at_lod_internal(bool overwrite_if_empty, uint32_t x, uint32_t y, uint32_t z, uint32_t insert_size,
                                   const svx_entry& e);
    int32_t clear_at_lod(uint32_t x, uint32_t y, uint32_t z, uint32_t clear_size);  // src/octree/update/clear.rs:55-348
    svx_entry get(uint32_t x, uint32_t y, uint32_t z) const;
    uint64_t structure_hash() const;

    // ---- MIP maps (host_octree_mip.cpp; reference src/octree/mipmap.rs, StrategyUpdater :716-938)
    bool mips_enabled() const { return mips_enabled_; }
    void switch_albedo_mip_maps(bool enabled);                                  // :858-872
    void recalculate_mips();                                                    // :798-855
    void mip_set_method_at(size_t level, uint32_t method, float thr);           // :657-672
    MipSampler mip_get_method_at(size_t level) const;                           // :650-655
    void mip_set_color_similarity_thr_at(size_t level, float thr);              // :617-630
    float mip_get_color_similarity_at(size_t level) const;                      // :610-615
    void mip_reset();                                                           // :718-721
    svx_entry sample_root_mip(uint32_t octant, uint32_t x, uint32_t y, uint32_t z) const;  // :897-937
    uint64_t mip_hash() const;
    const std::map<size_t, MipSampler>& mip_methods() const { return mip_methods_; }
    const std::map<size_t, float>& mip_thresholds() const { return mip_thresholds_; }
    // persistence hooks (host_octree_io.cpp)
    void mip_load_strategy(bool enabled, std::map<size_t, MipSampler> methods, std::map<size_t, float> thresholds);
    void mip_load_brick(size_t key, uint8_t kind, uint32_t solid, const uint32_t* voxels);
    // bencode persistence in the reference's byte format (host_octree_io.cpp; src/octree/mod.rs:138-168)
    uint8_t* to_bytes(size_t* len) const;  // malloc'd, the caller frees it with std::free; null = out of memory
    static int32_t from_bytes(const uint8_t* data, size_t len, HostOctree** out);
    int32_t save(const char* path) const;
    static int32_t load(const char* path, HostOctree** out);

    uint32_t size() const { return size_; }
    uint32_t brick_dim() const { return dim_; }
    uint32_t brick_volume() const { return vol_; }
    uint64_t revision() const { return revision_; }
    bool auto_simplify = true;

    // read access for the serialiser
    const std::vector<NodeRec>& nodes() const { return nodes_; }
    const uint32_t* brick_data(uint32_t handle) const { return voxels_.data() + (size_t)handle * vol_; }
    // the brick pool as the device mirrors it (handle-indexed) and the revision() at which a brick was last written:
    // a GPU copy taken at revision R is stale exactly in the bricks with brick_revision(h) > R
    size_t brick_pool_size() const { return brick_rev_.size(); }
    const uint32_t* brick_pool() const { return voxels_.data(); }
    uint64_t brick_revision(uint32_t handle) const { return brick_rev_[handle]; }
    const std::vector<svx_albedo>& color_palette() const { return colors_; }
    const std::vector<uint32_t>& data_palette() const { return datas_; }
    bool key_is_valid(size_t key) const { return key < nodes_.size() && nodes_[key].reserved; }
    uint64_t stored_occupied_bits(size_t key) const;
    bool value_is_empty(uint32_t v) const;  // pix_points_to_empty, src/octree/node.rs:405-427
    svx_entry resolve(uint32_t v) const;    // pix_get_ref, src/octree/node.rs:429-467

   private:
    // ---- arena (ObjectPool semantics, src/object_pool.rs:153-241) ----
    size_t pool_push();
    void pool_free(size_t key);
    bool next_available() const;
    // ---- brick pool ----
    uint32_t brick_alloc(uint32_t fill);
    uint32_t brick_clone(uint32_t handle);
    void brick_release(BrickRef& b);
    uint32_t* brick_mut(uint32_t handle) {  // any writer: what is known about the brick's 2x2x2 blocks is dropped
        brick_rev_[handle] = revision_;
        block_count_[handle] = BLOCKS_UNKNOWN;
        return voxels_.data() + (size_t)handle * vol_;
    }
    BrickRef brick_copy(const BrickRef& b);
    bool brick_equal(const BrickRef& a, const BrickRef& b) const;
    bool brick_homogeneous(const BrickRef& b, uint32_t* v) const;
    bool brick_simplify(BrickRef& b);
    bool brick_blockwise_uniform(uint32_t handle) const;
    uint64_t brick_bits(const uint32_t* vox) const;
    uint64_t brick_ref_bits(const BrickRef& b) const;
    void clear_content(size_t key);  // drops bricks, kind = Nothing
    // ---- reference algorithms ----
    uint32_t add_to_palette(const svx_entry& e);
    size_t leaf_update(bool overwrite, size_t key, const BoundsF& node_b, const BoundsF& target_b, size_t octant,
                       uint32_t x, uint32_t y, uint32_t z, uint32_t size, uint32_t content);
    size_t update_brick(bool overwrite, uint32_t handle, const BoundsF& b, uint32_t x, uint32_t y, uint32_t z,
                        uint32_t size, uint32_t data);
    void blocks_rebuild(uint32_t handle) const;
    void blocks_update(uint32_t handle, const size_t lo[3], const size_t hi[3]);
    void subdivide_leaf_to_nodes(size_t key, size_t target_octant);
    void deallocate_children_of(size_t key);
    void store_occupied_bits(size_t key, uint64_t bits);
    bool simplify(size_t key);
    bool node_is_all(const NodeRec& n, uint32_t v) const;
    bool node_compare(const NodeRec& a, const NodeRec& b) const;
    bool node_is_empty(const NodeRec& n) const;
    bool node_empty_at(size_t key, uint8_t octant) const;
    bool bitmap_cell_should_be_empty(size_t key, size_t x, size_t y, size_t z) const;
    bool brick_octant_empty(const BrickRef& b, uint8_t octant) const;
    bool brick_part_empty(const BrickRef& b, uint8_t part_octant, uint8_t target_octant) const;
    BrickRef try_brick_from_node(size_t key);
    void dilute(const uint32_t* src, uint32_t out_handles[8]);
    uint64_t hash_node(size_t key) const;
    uint64_t hash_brick(const BrickRef& b) const;
    uint64_t mip_hash_node(size_t key) const;
    svx_entry get_from(size_t key, BoundsF b, uint32_t x, uint32_t y, uint32_t z) const;  // get_internal, mod.rs:220-371
    void update_mip(size_t key, const BoundsF& nb, uint32_t x, uint32_t y, uint32_t z);  // mipmap.rs:296-584
    void recalculate_mip(size_t key, const BoundsF& nb);                                  // mipmap.rs:875-892
    bool mip_gather_children(size_t key, const uint32_t start[3], std::vector<svx_albedo>* out) const;
    bool mip_reduce(const MipSampler& how, const std::vector<svx_albedo>& samples, svx_albedo* out) const;
    void mip_defaults();

    uint32_t size_ = 0, dim_ = 0, vol_ = 0;
    std::vector<NodeRec> nodes_;
    size_t first_available_ = 0;
    std::vector<uint32_t> voxels_;
    std::vector<uint32_t> free_bricks_;
    std::vector<uint64_t> brick_rev_;  // per brick handle: revision_ of the mutation that last wrote it
    // Per brick: which aligned 2x2x2 blocks hold more than one value (one bit per block, index (bz * half + by) * half + bx)
    // and how many do - what the "8 bricks as one brick of half the resolution" test of simplify asks for after every insert
    // (brick_blockwise_uniform). Built on the first question, then kept exact by update_brick, the one in-place writer of the
    // edit path; every other writer goes through brick_mut, which drops the map (BLOCKS_UNKNOWN).
    static constexpr uint32_t BLOCKS_UNKNOWN = 0xFFFFFFFFu;
    mutable std::vector<uint64_t> block_bits_;
    mutable std::vector<uint32_t> block_count_;
    uint32_t block_words_ = 1;
    mutable std::vector<uint32_t> witness_;  // per brick: a voxel index known to differ from voxel 0 (0 = none known)
    std::vector<svx_albedo> colors_;
    std::vector<uint32_t> datas_;
    std::unordered_map<uint32_t, uint32_t> color_index_, data_index_;
    uint64_t revision_ = 0;
    bool mips_enabled_ = false;
    std::map<size_t, MipSampler> mip_methods_;
    std::map<size_t, float> mip_thresholds_;
    mutable std::vector<svx_albedo> mip_scratch_;
};

// Spatial helpers shared by host code (reference src/spatial/math/mod.rs)
uint64_t occupancy_box(uint32_t px, uint32_t py, uint32_t pz, uint32_t size, uint32_t dim);  // bits set_occupancy_in_bitmap_64bits would set

}  // namespace svx
