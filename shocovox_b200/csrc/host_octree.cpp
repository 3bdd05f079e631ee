// Host octree of the product (see host_octree.hpp). Construction follows the reference's rules so that the tree
// shape - and with it the f32 path of every ray - is the one the reference would build.
// `file:line` citations are relative to the reference checkout (shocovox-rs 0.11.1).
#include "host_octree.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace svx {

namespace {

constexpr uint32_t NONE16 = 0xFFFFu;

// Rust scalar casts: `f32 as usize` saturates, NaN and negatives give 0
inline size_t to_index(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}
// From<V3c<f32>> for V3c<usize> rounds first (src/spatial/math/vector.rs:306-316)
inline size_t round_index(float v) { return to_index(std::round(v)); }

inline size_t flat(size_t x, size_t y, size_t z, size_t dim) { return x + y * dim + z * dim * dim; }  // math/mod.rs:35

// octant <-> offset: x is bit 0, z bit 1, y bit 2 (hash_region, math/mod.rs:11-19; OCTANT_OFFSET_REGION_LUT, lut.rs:156)
inline float off_x(size_t o) { return (float)(o & 1); }
inline float off_y(size_t o) { return (float)((o >> 2) & 1); }
inline float off_z(size_t o) { return (float)((o >> 1) & 1); }

inline uint8_t octant_of(float dx, float dy, float dz, float half) {
    return (uint8_t)((dx >= half) + (dz >= half) * 2 + (dy >= half) * 4);
}

inline bool contains(const BoundsF& b, float x, float y, float z) {  // bound_contains, detail.rs:20-27
    return x >= b.x && x < b.x + b.size && y >= b.y && y < b.y + b.size && z >= b.z && z < b.z + b.size;
}

// matrix_index_for, math/mod.rs:44-77
inline void matrix_index(const BoundsF& b, uint32_t x, uint32_t y, uint32_t z, uint32_t dim, size_t out[3]) {
    out[0] = round_index(std::floor(((float)x - b.x) * (float)dim / b.size));
    out[1] = round_index(std::floor(((float)y - b.y) * (float)dim / b.size));
    out[2] = round_index(std::floor(((float)z - b.z) * (float)dim / b.size));
}

inline uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    return h;
}

}  // namespace

// set_occupancy_in_bitmap_64bits(position, size, dim, true, &0), math/mod.rs:114-162, returned as a mask
uint64_t occupancy_box(uint32_t px, uint32_t py, uint32_t pz, uint32_t size, uint32_t dim) {
    if (dim == 1) return ~0ull;
    const size_t count = to_index(std::ceil((float)size * 4.0f / (float)dim));
    const size_t sx = round_index(std::floor((float)(px * 4u) / (float)dim));
    const size_t sy = round_index(std::floor((float)(py * 4u) / (float)dim));
    const size_t sz = round_index(std::floor((float)(pz * 4u) / (float)dim));
    uint64_t m = 0;
    for (size_t x = sx; x < std::min<size_t>(sx + count, 4); ++x)
        for (size_t y = sy; y < std::min<size_t>(sy + count, 4); ++y)
            for (size_t z = sz; z < std::min<size_t>(sz + count, 4); ++z) m |= 1ull << (x + 4 * y + 16 * z);
    return m;
}

// ------------------------------------------------------------------------------------------------------------
// creation: Octree::new, src/octree/mod.rs:173-205
// ------------------------------------------------------------------------------------------------------------
int32_t HostOctree::create(uint32_t size, uint32_t brick_dim, HostOctree** out) {
    *out = nullptr;
    // `(v as f32).log(2.0).fract() != 0.0` : f32::log(self, base) = ln(self) / ln(base)
    auto not_whole_log2 = [](float v) {
        const float l = std::log(v) / std::log(2.0f);
        if (std::isinf(l) || std::isnan(l)) return true;  // fract() of +-inf is NaN, NaN != 0.0
        float ip;
        return std::modf(l, &ip) != 0.0f;
    };
    if (size == 0 || not_whole_log2((float)brick_dim)) return SVX_E_INVALID_BRICK_DIMENSION;
    if (brick_dim > size || not_whole_log2((float)size / (float)brick_dim)) return SVX_E_INVALID_SIZE;
    if (size < brick_dim * 2) return SVX_E_INVALID_STRUCTURE;
    HostOctree* t = new HostOctree();
    t->size_ = size;
    t->dim_ = brick_dim;
    t->vol_ = brick_dim * brick_dim * brick_dim;
    t->block_words_ = std::max<uint32_t>(1u, (t->vol_ / 8u + 63u) / 64u);  // one bit per 2x2x2 block
    t->pool_push();  // the root is key 0 and starts as Nothing
    t->mip_defaults();
    *out = t;
    return SVX_OK;
}

// ------------------------------------------------------------------------------------------------------------
// arena
// ------------------------------------------------------------------------------------------------------------
bool HostOctree::next_available() const {
    return first_available_ + 1 < nodes_.size() && !nodes_[first_available_ + 1].reserved;
}

size_t HostOctree::pool_push() {
    bool have;
    if (first_available_ < nodes_.size() && !nodes_[first_available_].reserved) {
        have = true;
    } else if (next_available()) {
        first_available_ += 1;
        have = true;
    } else {
        first_available_ = nodes_.size();
        have = false;
    }
    size_t key;
    if (have) {
        key = first_available_;
    } else {
        nodes_.emplace_back();
        key = nodes_.size() - 1;
    }
    nodes_[key].reserved = 1;
    if (next_available()) first_available_ += 1;
    clear_content(key);  // `*get_mut(key) = item`; the link part (node_children[key]) is NOT reset on reuse
    return key;
}

void HostOctree::pool_free(size_t key) {
    if (!key_is_valid(key)) return;
    clear_content(key);
    nodes_[key].reserved = 0;
    first_available_ = std::min(first_available_, key);
}

void HostOctree::clear_content(size_t key) {
    NodeRec& n = nodes_[key];
    for (auto& b : n.brick) brick_release(b);
    n.kind = NK_NOTHING;
    n.ocbits = 0;
}

// ------------------------------------------------------------------------------------------------------------
// brick pool
// ------------------------------------------------------------------------------------------------------------
uint32_t HostOctree::brick_alloc(uint32_t fill) {
    uint32_t h;
    if (!free_bricks_.empty()) {
        h = free_bricks_.back();
        free_bricks_.pop_back();
    } else {
        h = (uint32_t)(voxels_.size() / vol_);
        voxels_.resize(voxels_.size() + vol_);
        brick_rev_.push_back(revision_);
        witness_.push_back(0u);
        block_count_.push_back(BLOCKS_UNKNOWN);
        block_bits_.resize(block_bits_.size() + block_words_, 0ull);
    }
    std::fill_n(brick_mut(h), vol_, fill);
    witness_[h] = 0u;
    // one value everywhere: every block is uniform
    std::fill_n(block_bits_.begin() + (size_t)h * block_words_, block_words_, 0ull);
    block_count_[h] = 0u;
    return h;
}

uint32_t HostOctree::brick_clone(uint32_t handle) {
    const uint32_t h = brick_alloc(NIL);
    std::memcpy(brick_mut(h), brick_data(handle), (size_t)vol_ * 4);
    if (block_count_[handle] != BLOCKS_UNKNOWN) {  // the same voxels: the same blocks
        std::copy_n(block_bits_.begin() + (size_t)handle * block_words_, block_words_, block_bits_.begin() + (size_t)h * block_words_);
        block_count_[h] = block_count_[handle];
    }
    return h;
}

void HostOctree::brick_release(BrickRef& b) {
    if (b.kind == BK_PARTED) free_bricks_.push_back(b.value);
    b.kind = BK_EMPTY;
    b.value = NIL;
}

BrickRef HostOctree::brick_copy(const BrickRef& b) {
    BrickRef r = b;
    if (b.kind == BK_PARTED) r.value = brick_clone(b.value);
    return r;
}

bool HostOctree::brick_equal(const BrickRef& a, const BrickRef& b) const {  // derived PartialEq of BrickData
    if (a.kind != b.kind) return false;
    if (a.kind == BK_SOLID) return a.value == b.value;
    if (a.kind == BK_PARTED) return std::memcmp(brick_data(a.value), brick_data(b.value), (size_t)vol_ * 4) == 0;
    return true;
}

// get_homogeneous_data, src/octree/node.rs:300-313
bool HostOctree::brick_homogeneous(const BrickRef& b, uint32_t* v) const {
    if (b.kind == BK_EMPTY) return false;
    if (b.kind == BK_SOLID) {
        *v = b.value;
        return true;
    }
    // Same answer as the reference's linear scan, amortised: remember one position that differed from d[0] last
    // time (a "witness"); while it still differs the brick cannot be homogeneous.
    const uint32_t* d = brick_data(b.value);
    uint32_t& w = witness_[b.value];
    if (w != 0u && d[w] != d[0]) return false;
    for (uint32_t i = 1; i < vol_; ++i)
        if (d[i] != d[0]) {
            w = i;
            return false;
        }
    w = 0u;
    *v = d[0];
    return true;
}

// Is every aligned 2x2x2 block of the brick one value? (the test update/mod.rs:884-980 applies to Parted bricks when it
// tries to express 8 bricks as one brick of half the resolution - after EVERY insert that ends in a leaf). Same answer as
// the reference's scan, from the brick's block map (host_octree.hpp: block_bits_): a scan per question cost 2 us per
// voxel inserted into the 32^3 bricks of the blocky-terrain scene, most of its build time.
namespace {
inline bool block_is_uniform(const uint32_t* d, size_t bx, size_t by, size_t bz, uint32_t dim) {
    const uint32_t* p = d + flat(2 * bx, 2 * by, 2 * bz, dim);
    const size_t row = dim, plane = (size_t)dim * dim;
    const uint32_t v = p[0];
    return p[1] == v && p[row] == v && p[row + 1] == v && p[plane] == v && p[plane + 1] == v && p[plane + row] == v &&
           p[plane + row + 1] == v;
}
}  // namespace

void HostOctree::blocks_rebuild(uint32_t handle) const {
    const uint32_t* d = brick_data(handle);
    const size_t half = dim_ / 2;
    uint64_t* bits = block_bits_.data() + (size_t)handle * block_words_;
    std::fill_n(bits, block_words_, 0ull);
    uint32_t count = 0;
    size_t i = 0;
    for (size_t bz = 0; bz < half; ++bz)
        for (size_t by = 0; by < half; ++by)
            for (size_t bx = 0; bx < half; ++bx, ++i)
                if (!block_is_uniform(d, bx, by, bz, dim_)) {
                    bits[i >> 6] |= 1ull << (i & 63);
                    ++count;
                }
    block_count_[handle] = count;
}

// voxels [lo, hi) of the brick were just written in place: re-examine the blocks they touch
void HostOctree::blocks_update(uint32_t handle, const size_t lo[3], const size_t hi[3]) {
    if (block_count_[handle] == BLOCKS_UNKNOWN) return;
    const size_t half = dim_ / 2;
    if (half == 0 || lo[0] >= hi[0] || lo[1] >= hi[1] || lo[2] >= hi[2]) return;
    const uint32_t* d = brick_data(handle);
    uint64_t* bits = block_bits_.data() + (size_t)handle * block_words_;
    uint32_t count = block_count_[handle];
    for (size_t bz = lo[2] / 2; bz <= (hi[2] - 1) / 2; ++bz)
        for (size_t by = lo[1] / 2; by <= (hi[1] - 1) / 2; ++by)
            for (size_t bx = lo[0] / 2; bx <= (hi[0] - 1) / 2; ++bx) {
                const size_t i = (bz * half + by) * half + bx;
                const uint64_t bit = 1ull << (i & 63);
                const bool was = (bits[i >> 6] & bit) != 0, now = !block_is_uniform(d, bx, by, bz, dim_);
                if (was != now) {
                    bits[i >> 6] ^= bit;
                    count += now ? 1u : 0xFFFFFFFFu;  // +1 / -1
                }
            }
    block_count_[handle] = count;
}

bool HostOctree::brick_blockwise_uniform(uint32_t handle) const {
    if (dim_ < 2) return true;  // no blocks to disagree (the reference's loops are empty)
    if (block_count_[handle] == BLOCKS_UNKNOWN) blocks_rebuild(handle);
    return block_count_[handle] == 0u;
}

// BrickData::simplify, src/octree/node.rs:316-331
bool HostOctree::brick_simplify(BrickRef& b) {
    uint32_t v;
    if (!brick_homogeneous(b, &v)) return false;
    brick_release(b);
    if (!value_is_empty(v)) {
        b.kind = BK_SOLID;
        b.value = v;
    }
    return true;
}

// calculate_brick_occupied_bits, src/octree/node.rs:244-272
uint64_t HostOctree::brick_bits(const uint32_t* vox) const {
    uint64_t bits = 0;
    for (uint32_t z = 0; z < dim_; ++z)
        for (uint32_t y = 0; y < dim_; ++y)
            for (uint32_t x = 0; x < dim_; ++x)
                if (!value_is_empty(vox[flat(x, y, z, dim_)])) bits |= occupancy_box(x, y, z, 1, dim_);
    return bits;
}

// calculate_occupied_bits, src/octree/node.rs:275-297
uint64_t HostOctree::brick_ref_bits(const BrickRef& b) const {
    if (b.kind == BK_EMPTY) return 0;
    if (b.kind == BK_SOLID) return value_is_empty(b.value) ? 0 : ~0ull;
    return brick_bits(brick_data(b.value));
}

// ------------------------------------------------------------------------------------------------------------
// palette index values, src/octree/node.rs:354-467
// ------------------------------------------------------------------------------------------------------------
bool HostOctree::value_is_empty(uint32_t v) const {
    const uint32_t ci = v & 0xFFFFu, di = v >> 16;
    return (ci >= NONE16 || colors_[ci].a == 0) && (di == NONE16 || datas_[di] == 0);
}

svx_entry HostOctree::resolve(uint32_t v) const {
    svx_entry e{};
    const uint32_t ci = v & 0xFFFFu, di = v >> 16;
    const bool has_color = ci < NONE16, has_data = di != NONE16;
    if (!has_color && !has_data) return e;
    if (has_color) e.albedo = colors_[ci];
    if (has_data) e.data = datas_[di];
    e.kind = has_color && has_data ? SVX_ENTRY_COMPLEX : (has_color ? SVX_ENTRY_VISUAL : SVX_ENTRY_INFORMATIVE);
    return e;
}

// add_to_palette, src/octree/update/mod.rs:55-136
uint32_t HostOctree::add_to_palette(const svx_entry& e) {
    const svx_albedo a = e.albedo;
    const bool zero_albedo = (a.r | a.g | a.b | a.a) == 0;
    bool want_color = false, want_data = false;
    switch (e.kind) {
        case SVX_ENTRY_VISUAL: want_color = !zero_albedo; break;
        case SVX_ENTRY_INFORMATIVE: want_data = e.data != 0; break;
        case SVX_ENTRY_COMPLEX:
            want_color = !zero_albedo;
            want_data = e.data != 0;
            break;
        default: break;
    }
    uint32_t ci = NONE16, di = NONE16;
    if (want_color) {
        const uint32_t k = ((uint32_t)a.r << 24) | ((uint32_t)a.g << 16) | ((uint32_t)a.b << 8) | a.a;
        auto it = color_index_.find(k);
        if (it == color_index_.end()) {
            it = color_index_.emplace(k, (uint32_t)colors_.size()).first;
            colors_.push_back(a);
        }
        ci = it->second;
    }
    if (want_data) {
        auto it = data_index_.find(e.data);
        if (it == data_index_.end()) {
            it = data_index_.emplace(e.data, (uint32_t)datas_.size()).first;
            datas_.push_back(e.data);
        }
        di = it->second;
    }
    return (ci & 0xFFFFu) | (di << 16);
}

// ------------------------------------------------------------------------------------------------------------
// occupancy bits, src/octree/detail.rs:524-569
// ------------------------------------------------------------------------------------------------------------
uint64_t HostOctree::stored_occupied_bits(size_t key) const {
    const NodeRec& n = nodes_[key];
    if (n.kind == NK_INTERNAL) return n.ocbits;
    if (n.kind == NK_NOTHING) return 0;
    return n.link == LK_BITMAP ? n.leaf_bits : 0;
}

void HostOctree::store_occupied_bits(size_t key, uint64_t bits) {
    NodeRec& n = nodes_[key];
    if (n.kind == NK_INTERNAL) {
        n.ocbits = bits;
    } else {
        n.link = LK_BITMAP;  // Nothing / Leaf / UniformLeaf -> OccupancyBitmap(bits)
        n.leaf_bits = bits;
    }
}

// ------------------------------------------------------------------------------------------------------------
// get, src/octree/mod.rs:209-371
// ------------------------------------------------------------------------------------------------------------
svx_entry HostOctree::get(uint32_t x, uint32_t y, uint32_t z) const {
    return get_from(0, BoundsF{0, 0, 0, (float)size_}, x, y, z);
}

// get_internal, mod.rs:220-371: the same walk started at any node
svx_entry HostOctree::get_from(size_t key, BoundsF b, uint32_t x, uint32_t y, uint32_t z) const {
    const float px = (float)x, py = (float)y, pz = (float)z;
    svx_entry none{};
    if (!contains(b, px, py, pz)) return none;
    for (;;) {
        const NodeRec& n = nodes_[key];
        if (n.kind == NK_NOTHING) return none;
        const float half = b.size / 2.0f;
        const uint8_t oct = octant_of(px - b.x, py - b.y, pz - b.z, half);
        const BoundsF cb{b.x + off_x(oct) * half, b.y + off_y(oct) * half, b.z + off_z(oct) * half, half};
        if (n.kind == NK_INTERNAL) {
            const uint32_t c = n.link == LK_CHILDREN ? n.child[oct] : NIL;
            if (!key_is_valid(c)) return none;
            key = c;
            b = cb;
            continue;
        }
        const BrickRef& br = n.kind == NK_LEAF ? n.brick[oct] : n.brick[0];
        const BoundsF& bb = n.kind == NK_LEAF ? cb : b;
        if (br.kind == BK_EMPTY) return none;
        if (br.kind == BK_SOLID) {
            // a Leaf returns Solid values unconditionally (mod.rs:271-277), a UniformLeaf checks emptiness (:306-320)
            if (n.kind == NK_UNIFORM && value_is_empty(br.value)) return none;
            return resolve(br.value);
        }
        size_t mi[3];
        matrix_index(bb, x, y, z, dim_, mi);
        const uint32_t v = brick_data(br.value)[flat(mi[0], mi[1], mi[2], dim_)];
        if (value_is_empty(v)) return none;
        return resolve(v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// detail.rs helpers
// ------------------------------------------------------------------------------------------------------------
// try_brick_from_node, src/octree/detail.rs:489-500 (clones the brick)
BrickRef HostOctree::try_brick_from_node(size_t key) {
    if (!key_is_valid(key) || nodes_[key].kind != NK_UNIFORM) return BrickRef();
    return brick_copy(nodes_[key].brick[0]);
}

// deallocate_children_of, src/octree/detail.rs:503-520
void HostOctree::deallocate_children_of(size_t key) {
    if (!key_is_valid(key) || nodes_[key].link != LK_CHILDREN) return;
    uint32_t todo[8];
    int n = 0;
    for (int i = 0; i < 8; ++i)
        if (key_is_valid(nodes_[key].child[i])) todo[n++] = nodes_[key].child[i];
    for (int i = 0; i < n; ++i) {
        deallocate_children_of(todo[i]);
        pool_free(todo[i]);
        NodeRec& c = nodes_[todo[i]];
        c.link = LK_NONE;
        c.leaf_bits = 0;
        for (auto& k : c.child) k = NIL;
    }
}

// dilute_brick_data, src/octree/update/mod.rs:563-628 : 8 bricks, each a 2x magnification of one octant
void HostOctree::dilute(const uint32_t* src_in, uint32_t out[8]) {
    std::vector<uint32_t> src(src_in, src_in + vol_);  // the pool may grow while allocating
    for (size_t o = 0; o < 8; ++o) {
        if (dim_ == 1) {
            out[o] = brick_alloc(src[0]);
            continue;
        }
        if (dim_ == 2) {
            out[o] = brick_alloc(src[o]);  // (sic) indexes the source by octant number, update/mod.rs:582-595
            continue;
        }
        const size_t bx = (size_t)off_x(o) * 2, by = (size_t)off_y(o) * 2, bz = (size_t)off_z(o) * 2;
        const uint32_t h = brick_alloc(src[flat(bx, by, bz, dim_)]);
        uint32_t* dst = brick_mut(h);
        for (size_t x = 0; x < dim_; ++x)
            for (size_t y = 0; y < dim_; ++y)
                for (size_t z = 0; z < dim_; ++z) {
                    if (x < 2 && y < 2 && z < 2) continue;
                    dst[flat(x, y, z, dim_)] = src[flat(bx + x / 2, by + y / 2, bz + z / 2, dim_)];
                }
        out[o] = h;
    }
}

// subdivide_leaf_to_nodes, src/octree/detail.rs:321-486
void HostOctree::subdivide_leaf_to_nodes(size_t key, size_t target_octant) {
    // take the leaf content out; the node becomes Internal(occupancy bitmap)
    const uint8_t old_kind = nodes_[key].kind;
    BrickRef old[8];
    for (int i = 0; i < 8; ++i) {
        old[i] = nodes_[key].brick[i];
        nodes_[key].brick[i] = BrickRef();
    }
    nodes_[key].kind = NK_INTERNAL;
    nodes_[key].ocbits = nodes_[key].leaf_bits;

    uint32_t fresh[8] = {NIL, NIL, NIL, NIL, NIL, NIL, NIL, NIL};
    auto push_uniform = [&](const BrickRef& b, uint64_t bits) {
        const size_t c = pool_push();
        NodeRec& n = nodes_[c];
        n.kind = NK_UNIFORM;
        n.brick[0] = b;
        n.link = LK_BITMAP;
        n.leaf_bits = bits;
        return (uint32_t)c;
    };
    if (old_kind == NK_LEAF) {
        for (size_t o = 0; o < 8; ++o) {
            if (old[o].kind == BK_EMPTY) {
                if (o == target_octant) fresh[o] = (uint32_t)pool_push();  // an empty child; its link is untouched
            } else if (old[o].kind == BK_SOLID) {
                fresh[o] = push_uniform(old[o], ~0ull);
            } else {
                // detail.rs:403-410 computes the bits from `bricks[octant]`, which :343 already swapped to Empty -> 0
                fresh[o] = push_uniform(old[o], 0);
            }
        }
    } else if (old_kind == NK_UNIFORM) {
        BrickRef& b = old[0];
        if (b.kind == BK_EMPTY) {
            const size_t c = pool_push();
            nodes_[c].link = LK_BITMAP;
            nodes_[c].leaf_bits = 0;
            fresh[target_octant] = (uint32_t)c;
        } else if (b.kind == BK_SOLID) {
            for (size_t o = 0; o < 8; ++o) fresh[o] = push_uniform(b, ~0ull);
        } else {
            uint32_t parts[8];
            dilute(brick_data(b.value), parts);
            for (size_t o = 0; o < 8; ++o) {
                const uint64_t bits = brick_bits(brick_data(parts[o]));
                BrickRef nb;
                nb.kind = BK_PARTED;
                nb.value = parts[o];
                fresh[o] = push_uniform(nb, bits);
            }
            brick_release(b);
        }
    }
    NodeRec& n = nodes_[key];
    n.link = LK_CHILDREN;
    for (int i = 0; i < 8; ++i) n.child[i] = fresh[i];
}

// ------------------------------------------------------------------------------------------------------------
// update_brick, src/octree/update/mod.rs:637-675
// ------------------------------------------------------------------------------------------------------------
size_t HostOctree::update_brick(bool overwrite, uint32_t handle, const BoundsF& b, uint32_t x, uint32_t y, uint32_t z,
                                uint32_t size, uint32_t data) {
    size_t mi[3];
    matrix_index(b, x, y, z, dim_, mi);
    const size_t update_size = std::min<size_t>((size_t)dim_ - mi[0], size);
    const bool color_some = (data & 0xFFFFu) < NONE16, data_some = (data >> 16) != NONE16;
    // written in place: the brick's revision moves, its block map is brought up to date below instead of being dropped
    brick_rev_[handle] = revision_;
    uint32_t* brick = voxels_.data() + (size_t)handle * vol_;
    const size_t hi[3] = {std::min<size_t>(mi[0] + size, dim_), std::min<size_t>(mi[1] + size, dim_), std::min<size_t>(mi[2] + size, dim_)};
    // the same box of voxels as the reference's x / y / z loops (update/mod.rs:637-675), walked in memory order (x fastest)
    for (size_t iz = mi[2]; iz < hi[2]; ++iz)
        for (size_t iy = mi[1]; iy < hi[1]; ++iy)
            for (size_t ix = mi[0]; ix < hi[0]; ++ix) {
                uint32_t& v = brick[flat(ix, iy, iz, dim_)];
                if (overwrite) {
                    v = data;
                } else {
                    if (color_some) v = (v & 0xFFFF0000u) | (data & 0x0000FFFFu);
                    if (data_some) v = (v & 0x0000FFFFu) | (data & 0xFFFF0000u);
                }
            }
    blocks_update(handle, mi, hi);
    return update_size;
}

// leaf_update, src/octree/update/mod.rs:160-549
size_t HostOctree::leaf_update(bool overwrite, size_t key, const BoundsF& node_b, const BoundsF& target_b, size_t octant,
                               uint32_t x, uint32_t y, uint32_t z, uint32_t size, uint32_t content) {
    const bool content_empty = value_is_empty(content);
    NodeRec& n = nodes_[key];
    if (n.kind == NK_LEAF) {
        BrickRef& b = n.brick[octant];
        if (b.kind == BK_EMPTY) {
            const uint32_t h = brick_alloc(NIL);
            const size_t us = update_brick(overwrite, h, target_b, x, y, z, size, content);
            BrickRef& nb = nodes_[key].brick[octant];
            nb.kind = BK_PARTED;
            nb.value = h;
            return us;
        }
        if (b.kind == BK_SOLID) {
            const uint32_t voxel = b.value;
            if ((content_empty && !value_is_empty(voxel)) || (!content_empty && voxel != content)) {
                const uint32_t h = brick_alloc(voxel);
                const size_t us = update_brick(overwrite, h, target_b, x, y, z, size, content);
                BrickRef& nb = nodes_[key].brick[octant];
                nb.kind = BK_PARTED;
                nb.value = h;
                return us;
            }
            return 0;
        }
        return update_brick(overwrite, b.value, target_b, x, y, z, size, content);
    }
    if (n.kind == NK_UNIFORM) {
        BrickRef& mat = n.brick[0];
        if (mat.kind == BK_EMPTY) {
            if (content_empty) return 0;  // the reference recurses forever here; insert() never reaches it (insert.rs:117)
            const uint32_t h = brick_alloc(NIL);
            const size_t us = update_brick(overwrite, h, target_b, x, y, z, size, content);
            NodeRec& m = nodes_[key];
            clear_content(key);
            m.kind = NK_LEAF;
            m.brick[octant].kind = BK_PARTED;
            m.brick[octant].value = h;
            return us;
        }
        if (mat.kind == BK_SOLID) {
            const uint32_t voxel = mat.value;
            const bool voxel_empty = value_is_empty(voxel);
            if (content_empty && voxel_empty) {
                clear_content(key);
                return 0;
            }
            if ((!content_empty && voxel != content) || (content_empty && !voxel_empty)) {
                const uint32_t h = brick_alloc(voxel);
                BrickRef& m = nodes_[key].brick[0];
                m.kind = BK_PARTED;
                m.value = h;
                return leaf_update(overwrite, key, node_b, target_b, octant, x, y, z, size, content);
            }
            return 0;
        }
        // Parted uniform leaf
        size_t mi[3];
        matrix_index(node_b, x, y, z, dim_, mi);
        const uint32_t cur = brick_data(mat.value)[flat(mi[0], mi[1], mi[2], dim_)];
        if (dim_ > 1 && ((content_empty && value_is_empty(cur)) || (!content_empty && cur == content))) return 0;
        if (node_b.size <= (float)dim_ && dim_ > 1)
            return update_brick(overwrite, mat.value, node_b, x, y, z, size, content);
        // split the uniform leaf into 8 bricks
        uint32_t parts[8];
        size_t us = 0;
        const uint32_t src_handle = mat.value;
        if (dim_ == 1) {
            for (size_t o = 0; o < 8; ++o) parts[o] = brick_clone(src_handle);
            us = update_brick(overwrite, parts[octant], target_b, x, y, z, size, content);
        } else {
            dilute(brick_data(src_handle), parts);
            us = update_brick(overwrite, parts[octant], target_b, x, y, z, size, content);
        }
        clear_content(key);  // releases the source brick
        NodeRec& m = nodes_[key];
        m.kind = NK_LEAF;
        for (size_t o = 0; o < 8; ++o) {
            m.brick[o].kind = BK_PARTED;
            m.brick[o].value = parts[o];
        }
        return us;
    }
    // Internal / Nothing: gather what the children hold as bricks, drop the children (update/mod.rs:497-547).
    // For Internal nodes the link is switched to the bitmap FIRST, so `child(o)` already reads as empty and the
    // children are neither harvested nor freed ("might induce data loss - see #69").
    if (n.kind == NK_INTERNAL) {
        n.link = LK_BITMAP;
        n.leaf_bits = n.ocbits;
    }
    BrickRef gathered[8];
    for (size_t o = 0; o < 8; ++o) {
        const uint32_t c = nodes_[key].link == LK_CHILDREN ? nodes_[key].child[o] : NIL;
        gathered[o] = try_brick_from_node(c);
    }
    clear_content(key);
    for (size_t o = 0; o < 8; ++o) nodes_[key].brick[o] = gathered[o];
    nodes_[key].kind = NK_LEAF;
    deallocate_children_of(key);
    return leaf_update(overwrite, key, node_b, target_b, octant, x, y, z, size, content);
}

// ------------------------------------------------------------------------------------------------------------
// simplify, src/octree/update/mod.rs:689-1068
// ------------------------------------------------------------------------------------------------------------
bool HostOctree::node_is_all(const NodeRec& n, uint32_t v) const {  // node.rs:518-552
    auto all = [&](const BrickRef& b) {
        uint32_t h;
        return brick_homogeneous(b, &h) && h == v;
    };
    if (n.kind == NK_UNIFORM) return all(n.brick[0]);
    if (n.kind == NK_LEAF) {
        for (int o = 0; o < 8; ++o)
            if (!all(n.brick[o])) return false;
        return true;
    }
    return false;
}

bool HostOctree::node_compare(const NodeRec& a, const NodeRec& b) const {  // node.rs:554-573
    if (a.kind == NK_NOTHING) return b.kind == NK_NOTHING;
    if (a.kind == NK_INTERNAL || a.kind != b.kind) return false;
    if (a.kind == NK_UNIFORM) return brick_equal(a.brick[0], b.brick[0]);
    for (int o = 0; o < 8; ++o)
        if (!brick_equal(a.brick[o], b.brick[o])) return false;
    return true;
}

bool HostOctree::simplify(size_t key) {
    if (!key_is_valid(key)) return false;
    NodeRec& n = nodes_[key];
    switch (n.kind) {
        case NK_NOTHING: return true;
        case NK_UNIFORM: {
            BrickRef& b = n.brick[0];
            if (b.kind == BK_EMPTY) return true;
            if (b.kind == BK_SOLID) {
                if (!value_is_empty(b.value)) return false;
                clear_content(key);
                n.link = LK_NONE;
                n.leaf_bits = 0;
                for (auto& k : n.child) k = NIL;
                return true;
            }
            return brick_simplify(b);
        }
        case NK_LEAF: {
            bool simplified = false, uniform_solid = true, have = false;
            uint32_t solid_value = 0;
            for (int o = 0; o < 8; ++o) {
                simplified |= brick_simplify(n.brick[o]);
                if (uniform_solid) {
                    if (n.brick[o].kind == BK_SOLID) {
                        if (have) {
                            if (solid_value != n.brick[o].value) uniform_solid = false;
                        } else {
                            have = true;
                            solid_value = n.brick[o].value;
                        }
                    } else {
                        uniform_solid = false;
                    }
                }
            }
            if (uniform_solid) {
                clear_content(key);
                n.kind = NK_UNIFORM;
                n.brick[0].kind = BK_SOLID;
                n.brick[0].value = solid_value;
                return true;
            }
            // can the 8 bricks be expressed as ONE brick at half resolution?
            const size_t half = dim_ / 2;
            bool uniform = true;
            for (int o = 0; o < 8 && uniform; ++o) {
                const BrickRef& b = n.brick[o];
                if (b.kind != BK_PARTED) {
                    uniform = brick_equal(b, n.brick[0]);
                    continue;
                }
                uniform = brick_blockwise_uniform(b.value);
            }
            if (!uniform) return simplified;
            const uint32_t h = brick_alloc(NIL);
            for (size_t o = 0; o < 8; ++o) {
                const BrickRef b = nodes_[key].brick[o];
                if (b.kind == BK_EMPTY) continue;
                uint32_t* dst = brick_mut(h);
                const size_t ox = round_index(off_x(o) * (float)half), oy = round_index(off_y(o) * (float)half),
                             oz = round_index(off_z(o) * (float)half);
                for (size_t x = 0; x < half; ++x)
                    for (size_t y = 0; y < half; ++y)
                        for (size_t z = 0; z < half; ++z)
                            dst[flat(ox + x, oy + y, oz + z, dim_)] =
                                b.kind == BK_SOLID ? b.value : brick_data(b.value)[flat(2 * x, 2 * y, 2 * z, dim_)];
            }
            clear_content(key);
            NodeRec& m = nodes_[key];
            m.kind = NK_UNIFORM;
            m.brick[0].kind = BK_PARTED;
            m.brick[0].value = h;
            return true;
        }
        case NK_INTERNAL: {
            if (n.ocbits == 0 || n.link == LK_NONE) {
                clear_content(key);
                return true;
            }
            if (n.link != LK_CHILDREN) return false;
            uint32_t kids[8];
            for (int i = 0; i < 8; ++i) kids[i] = n.child[i];
            simplify(kids[0]);
            if (!key_is_valid(kids[0])) {
                for (int i = 1; i < 8; ++i) simplify(kids[i]);
                return false;
            }
            for (int o = 1; o < 8; ++o) {
                simplify(kids[o]);
                if (!key_is_valid(kids[o]) || !node_compare(nodes_[kids[0]], nodes_[kids[o]])) return false;
            }
            // all 8 children hold the same leaf: this node becomes that leaf (update/mod.rs:1046-1061)
            NodeRec& parent = nodes_[key];
            NodeRec& first = nodes_[kids[0]];
            std::swap(parent.kind, first.kind);
            std::swap(parent.ocbits, first.ocbits);
            for (int i = 0; i < 8; ++i) std::swap(parent.brick[i], first.brick[i]);
            const uint8_t link = first.link;
            const uint64_t bits = first.leaf_bits;
            uint32_t link_children[8];
            for (int i = 0; i < 8; ++i) link_children[i] = first.child[i];
            deallocate_children_of(key);
            NodeRec& p = nodes_[key];
            p.link = link;
            p.leaf_bits = bits;
            for (int i = 0; i < 8; ++i) p.child[i] = link_children[i];
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------------------
// insert_at_lod_internal, src/octree/update/insert.rs:99-388
// ------------------------------------------------------------------------------------------------------------
int32_t HostOctree::insert_at_lod_internal(bool overwrite, uint32_t x, uint32_t y, uint32_t z, uint32_t insert_size,
                                           const svx_entry& e) {
    const BoundsF root{0, 0, 0, (float)size_};
    const float px = (float)x, py = (float)y, pz = (float)z;
    if (!contains(root, px, py, pz)) return SVX_E_INVALID_POSITION;
    // OctreeEntry::is_none, src/octree/mod.rs:102-109
    const bool no_color = e.albedo.a == 0, no_data = e.data == 0;
    if (e.kind == SVX_ENTRY_EMPTY || (e.kind == SVX_ENTRY_VISUAL && no_color) ||
        (e.kind == SVX_ENTRY_INFORMATIVE && no_data) || (e.kind == SVX_ENTRY_COMPLEX && no_color && no_data))
        return SVX_OK;
    if (e.kind > SVX_ENTRY_COMPLEX) return SVX_E_INVALID_ARGUMENT;
    ++revision_;

    struct Visit {
        uint32_t key;
        BoundsF b;
    };
    Visit path[64];
    int depth = 0;
    path[depth++] = {0u, root};
    size_t actual_update_size = 0;
    const uint32_t content = add_to_palette(e);

    for (;;) {
        const size_t cur = path[depth - 1].key;
        const BoundsF cb = path[depth - 1].b;
        const uint8_t oct = octant_of(px - cb.x, py - cb.y, pz - cb.z, cb.size / 2.0f);
        const BoundsF tb{cb.x + off_x(oct) * cb.size / 2.0f, cb.y + off_y(oct) * cb.size / 2.0f,
                         cb.z + off_z(oct) * cb.size / 2.0f, cb.size / 2.0f};
        uint32_t child = nodes_[cur].link == LK_CHILDREN ? nodes_[cur].child[oct] : NIL;

        // lexicographic `position <= target_bounds.min_position` (derived PartialOrd, vector.rs:3)
        const bool pos_le_min = px != tb.x ? px < tb.x : (py != tb.y ? py < tb.y : pz <= tb.z);
        if (insert_size > 1 && tb.size <= (float)insert_size && pos_le_min) {
            // the whole child is overwritten with one solid value
            if (nodes_[cur].kind == NK_LEAF || nodes_[cur].kind == NK_UNIFORM) {
                subdivide_leaf_to_nodes(cur, oct);
                child = nodes_[cur].link == LK_CHILDREN ? nodes_[cur].child[oct] : NIL;
            }
            size_t target;
            if (key_is_valid(child)) {
                deallocate_children_of(child);
                clear_content(child);
                target = child;
            } else {
                target = pool_push();
                NodeRec& p = nodes_[cur];
                if (p.link == LK_NONE) {  // child_mut, node.rs:56-64
                    p.link = LK_CHILDREN;
                    for (auto& k : p.child) k = NIL;
                }
                p.child[oct] = (uint32_t)target;
            }
            NodeRec& t = nodes_[target];
            t.kind = NK_UNIFORM;
            t.brick[0].kind = BK_SOLID;
            t.brick[0].value = content;
            t.link = LK_BITMAP;
            t.leaf_bits = ~0ull;
            for (auto& k : t.child) k = NIL;
            actual_update_size = to_index(tb.size);
            break;
        }

        const uint8_t kind = nodes_[cur].kind;
        const float data_size = kind == NK_UNIFORM ? cb.size : tb.size;
        if (data_size > (float)dim_ || key_is_valid(child)) {
            if (key_is_valid(child)) {
                path[depth++] = {child, tb};
            } else if (kind == NK_LEAF || kind == NK_UNIFORM) {
                const NodeRec& n = nodes_[cur];
                const BrickRef& br = kind == NK_UNIFORM ? n.brick[0] : n.brick[oct];
                bool match = false;
                if (br.kind == BK_SOLID) {
                    match = br.value == content;
                } else if (br.kind == BK_PARTED) {
                    size_t mi[3];
                    matrix_index(kind == NK_UNIFORM ? cb : tb, x, y, z, dim_, mi);
                    match = brick_data(br.value)[flat(mi[0], mi[1], mi[2], dim_)] == content;
                }
                if (match || node_is_all(n, content)) break;
                subdivide_leaf_to_nodes(cur, oct);
                path[depth++] = {nodes_[cur].child[oct], tb};
            } else {
                if (kind == NK_NOTHING) {
                    nodes_[cur].kind = NK_INTERNAL;
                    nodes_[cur].ocbits = 0;
                }
                const size_t fresh = pool_push();
                NodeRec& p = nodes_[cur];
                if (p.link == LK_NONE) {
                    p.link = LK_CHILDREN;
                    for (auto& k : p.child) k = NIL;
                }
                p.child[oct] = (uint32_t)fresh;
                path[depth++] = {(uint32_t)fresh, tb};
            }
        } else {
            actual_update_size = leaf_update(overwrite, cur, cb, tb, oct, x, y, z, insert_size, content);
            break;
        }
    }

    // post-processing, insert.rs:317-386 : occupancy bits bottom-up, then simplification
    bool simplifyable = auto_simplify;
    for (int i = depth - 1; i >= 0; --i) {
        const size_t key = path[i].key;
        const BoundsF nb = path[i].b;
        if (!key_is_valid(key)) continue;
        if (nodes_[key].kind == NK_NOTHING) {
            nodes_[key].kind = NK_INTERNAL;
            nodes_[key].ocbits = 0;
        }
        uint64_t bits = stored_occupied_bits(key);
        if (to_index(nb.size) == actual_update_size) {
            bits = ~0ull;
        } else {
            bits |= occupancy_box((uint32_t)round_index(px - nb.x), (uint32_t)round_index(py - nb.y),
                                  (uint32_t)round_index(pz - nb.z), (uint32_t)actual_update_size, (uint32_t)to_index(nb.size));
        }
        store_occupied_bits(key, bits);
        update_mip(key, nb, x, y, z);  // insert.rs:371
        const uint8_t k = nodes_[key].kind;
        if (k == NK_LEAF || k == NK_UNIFORM) {
            simplifyable = simplify(key);
            continue;
        }
        if (simplifyable) simplifyable = simplify(key);
    }
    return SVX_OK;
}

// ------------------------------------------------------------------------------------------------------------
// clear / clear_at_lod, src/octree/update/clear.rs:48-348, with the emptiness predicates of node.rs / detail.rs
// ------------------------------------------------------------------------------------------------------------
// BrickData::is_empty_throughout, node.rs:107-178: is one octant of the brick empty?
bool HostOctree::brick_octant_empty(const BrickRef& b, uint8_t octant) const {
    if (b.kind == BK_EMPTY) return true;
    if (b.kind == BK_SOLID) return value_is_empty(b.value);
    const uint32_t* d = brick_data(b.value);
    if (dim_ == 1) return value_is_empty(d[0]);
    const size_t ox = octant & 1, oy = (octant >> 2) & 1, oz = (octant >> 1) & 1;
    if (dim_ == 2) return value_is_empty(d[flat(ox, oy, oz, 2)]);
    const size_t e = dim_ / 2;
    for (size_t x = ox * e; x < ox * e + e; ++x)
        for (size_t y = oy * e; y < oy * e + e; ++y)
            for (size_t z = oz * e; z < oz * e + e; ++z)
                if (!value_is_empty(d[flat(x, y, z, dim_)])) return false;
    return true;
}

// BrickData::is_part_empty_throughout, node.rs:184-241: is sub-octant `target` of octant `part` empty?
bool HostOctree::brick_part_empty(const BrickRef& b, uint8_t part, uint8_t target) const {
    if (b.kind == BK_EMPTY) return true;
    if (b.kind == BK_SOLID) return value_is_empty(b.value);
    const uint32_t* d = brick_data(b.value);
    if (dim_ == 1) return value_is_empty(d[0]);
    if (dim_ == 2) return value_is_empty(d[flat(part & 1, (part >> 2) & 1, (part >> 1) & 1, 2)]);
    const float outer = (float)dim_ / 2.0f, inner = (float)dim_ / 4.0f;
    const size_t bx = round_index(off_x(part) * outer + off_x(target) * inner), by = round_index(off_y(part) * outer + off_y(target) * inner),
                 bz = round_index(off_z(part) * outer + off_z(target) * inner);
    const size_t n = to_index(inner);
    for (size_t x = 0; x < n; ++x)
        for (size_t y = 0; y < n; ++y)
            for (size_t z = 0; z < n; ++z)
                if (!value_is_empty(d[flat(bx + x, by + y, bz + z, dim_)])) return false;
    return true;
}

// NodeContent::is_empty, node.rs:470-515
bool HostOctree::node_is_empty(const NodeRec& n) const {
    auto empty = [&](const BrickRef& b) {
        if (b.kind == BK_EMPTY) return true;
        if (b.kind == BK_SOLID) return value_is_empty(b.value);
        const uint32_t* d = brick_data(b.value);
        for (uint32_t i = 0; i < vol_; ++i)
            if (!value_is_empty(d[i])) return false;
        return true;
    };
    if (n.kind == NK_NOTHING) return true;
    if (n.kind == NK_INTERNAL) return false;
    if (n.kind == NK_UNIFORM) return empty(n.brick[0]);
    for (int o = 0; o < 8; ++o)
        if (!empty(n.brick[o])) return false;
    return true;
}

// node_empty_at, detail.rs:255-316
bool HostOctree::node_empty_at(size_t key, uint8_t octant) const {
    const NodeRec& n = nodes_[key];
    auto empty = [&](const BrickRef& b) {
        if (b.kind == BK_EMPTY) return true;
        uint32_t v;
        return brick_homogeneous(b, &v) ? value_is_empty(v) : false;
    };
    switch (n.kind) {
        case NK_NOTHING: return true;
        case NK_LEAF: return empty(n.brick[octant]);
        case NK_UNIFORM: return empty(n.brick[0]);
        default: {
            // (sic) looks at the child under `octant` and asks it about each of ITS octants (detail.rs:305-313)
            const uint32_t c = n.link == LK_CHILDREN ? n.child[octant] : NIL;
            if (!key_is_valid(c)) return true;
            for (uint8_t o = 0; o < 8; ++o)
                if (!node_empty_at(c, o)) return false;
            return true;
        }
    }
}

// should_bitmap_be_empty_at_bitmap_index, detail.rs:181-252
bool HostOctree::bitmap_cell_should_be_empty(size_t key, size_t x, size_t y, size_t z) const {
    const float px = 0.5f + (float)x, py = 0.5f + (float)y, pz = 0.5f + (float)z;
    const uint8_t oct = octant_of(px, py, pz, 2.0f);
    const uint8_t sub = octant_of(px - off_x(oct) * 4.0f / 2.0f, py - off_y(oct) * 4.0f / 2.0f, pz - off_z(oct) * 4.0f / 2.0f, 1.0f);
    const NodeRec& n = nodes_[key];
    switch (n.kind) {
        case NK_NOTHING: return true;
        case NK_INTERNAL: {
            const uint32_t c = n.link == LK_CHILDREN ? n.child[oct] : NIL;
            return key_is_valid(c) ? node_empty_at(c, sub) : true;
        }
        case NK_UNIFORM: return brick_part_empty(n.brick[0], oct, sub);
        default: return brick_octant_empty(n.brick[oct], sub);
    }
}

int32_t HostOctree::clear_at_lod(uint32_t x, uint32_t y, uint32_t z, uint32_t clear_size) {
    const BoundsF root{0, 0, 0, (float)size_};
    const float px = (float)x, py = (float)y, pz = (float)z;
    if (!contains(root, px, py, pz)) return SVX_E_INVALID_POSITION;
    ++revision_;
    struct Visit {
        uint32_t key;
        BoundsF b;
    };
    Visit path[64];
    int depth = 0;
    path[depth++] = {0u, root};
    size_t actual_update_size = 0;
    auto as_u32 = [](float v) { return (uint32_t)round_index(v); };

    for (;;) {
        const size_t cur = path[depth - 1].key;
        const BoundsF cb = path[depth - 1].b;
        const uint8_t oct = octant_of(px - cb.x, py - cb.y, pz - cb.z, cb.size / 2.0f);
        const BoundsF tb{cb.x + off_x(oct) * cb.size / 2.0f, cb.y + off_y(oct) * cb.size / 2.0f,
                         cb.z + off_z(oct) * cb.size / 2.0f, cb.size / 2.0f};
        const uint32_t child = nodes_[cur].link == LK_CHILDREN ? nodes_[cur].child[oct] : NIL;
        const uint32_t mx = as_u32(tb.x), my = as_u32(tb.y), mz = as_u32(tb.z);
        const bool pos_le_min = x != mx ? x < mx : (y != my ? y < my : z <= mz);  // V3c<u32> lexicographic <=
        if (clear_size > 1 && tb.size <= (float)clear_size && pos_le_min && key_is_valid(child)) {
            // the whole child node is erased; the parents' occupancy is repaired below
            deallocate_children_of(child);
            clear_content(child);
            NodeRec& c = nodes_[child];
            c.link = LK_NONE;
            c.leaf_bits = 0;
            for (auto& k : c.child) k = NIL;
            actual_update_size = to_index(tb.size);
            path[depth++] = {child, tb};
            break;
        }
        if (tb.size > (float)std::max(clear_size, dim_) || key_is_valid(child)) {
            if (key_is_valid(child)) {
                path[depth++] = {child, tb};
                continue;
            }
            const NodeRec& n = nodes_[cur];
            if (n.kind != NK_LEAF && n.kind != NK_UNIFORM) break;  // nothing stored here
            const BrickRef& br = n.kind == NK_UNIFORM ? n.brick[0] : n.brick[oct];
            bool match;
            if (br.kind == BK_EMPTY) {
                match = true;
            } else if (br.kind == BK_SOLID) {
                match = value_is_empty(br.value);
            } else {
                // (sic) unscaled `position - current_bounds.min` as brick index (clear.rs:139-146); the reference
                // bounds-panics when it leaves the brick
                const size_t fi = flat(x - as_u32(cb.x), y - as_u32(cb.y), z - as_u32(cb.z), dim_);
                if (fi >= vol_) return SVX_E_INVALID_STRUCTURE;
                match = value_is_empty(brick_data(br.value)[fi]);
            }
            if (match || node_is_empty(n)) break;
            subdivide_leaf_to_nodes(cur, oct);
            path[depth++] = {nodes_[cur].link == LK_CHILDREN ? nodes_[cur].child[oct] : NIL, tb};
        } else {
            actual_update_size = leaf_update(true, cur, cb, tb, oct, x, y, z, clear_size, NIL);
            break;
        }
    }

    // post-processing, clear.rs:224-346: free removed nodes, recompute the occupancy bits of the ancestors
    bool have_removed = false;
    Visit removed = path[depth - 1];
    --depth;
    if (to_index(removed.b.size) <= actual_update_size) have_removed = true;
    bool simplifyable = auto_simplify;
    for (int i = depth - 1; i >= 0; --i) {
        const size_t key = path[i].key;
        const BoundsF nb = path[i].b;
        if (have_removed) {
            const uint8_t co = octant_of((removed.b.x - nb.x) + removed.b.size / 2.0f, (removed.b.y - nb.y) + removed.b.size / 2.0f,
                                         (removed.b.z - nb.z) + removed.b.size / 2.0f, nb.size / 2.0f);
            NodeRec& n = nodes_[key];
            if (n.link == LK_CHILDREN) {  // NodeChildren::clear, node.rs:75-83
                n.child[co] = NIL;
                bool none = true;
                for (int k = 0; k < 8; ++k) none &= n.child[k] == NIL;
                if (none) {
                    n.link = LK_NONE;
                    n.leaf_bits = 0;
                }
            }
            pool_free(removed.key);
            have_removed = false;
        }
        const uint64_t previous = stored_occupied_bits(key);
        uint64_t bits = nodes_[key].link == LK_NONE ? 0 : previous;
        if (to_index(nb.size) == actual_update_size) {
            bits = 0;
        } else {
            size_t s[3];
            matrix_index(nb, x, y, z, 4, s);
            const size_t n = to_index(std::ceil((float)actual_update_size * 4.0f / nb.size));
            for (size_t cx = s[0]; cx < std::min<size_t>(s[0] + n, 4); ++cx)
                for (size_t cy = s[1]; cy < std::min<size_t>(s[1] + n, 4); ++cy)
                    for (size_t cz = s[2]; cz < std::min<size_t>(s[2] + n, 4); ++cz)
                        if (bitmap_cell_should_be_empty(key, cx, cy, cz)) bits &= ~(1ull << (cx + 4 * cy + 16 * cz));
        }
        if (bits != 0 && nodes_[key].link == LK_CHILDREN) {
            clear_content(key);
            nodes_[key].kind = NK_INTERNAL;
            nodes_[key].ocbits = bits;
        } else {
            deallocate_children_of(key);
            NodeRec& n = nodes_[key];
            n.link = LK_NONE;
            n.leaf_bits = 0;
            for (auto& k : n.child) k = NIL;
            have_removed = true;
            removed = path[i];
            clear_content(key);
        }
        if (bits == 0) {
            NodeRec& n = nodes_[key];
            n.link = LK_NONE;
            n.leaf_bits = 0;
            for (auto& k : n.child) k = NIL;
        } else {
            store_occupied_bits(key, bits);
        }
        update_mip(key, path[i].b, x, y, z);  // clear.rs:335
        if (simplifyable) simplifyable = simplify(key);
        if (previous == bits) break;
    }
    return SVX_OK;
}

// ------------------------------------------------------------------------------------------------------------
// structure hash (same definition as the oracle's: what a ray can observe, independent of key numbering)
// ------------------------------------------------------------------------------------------------------------
uint64_t HostOctree::hash_brick(const BrickRef& b) const {
    uint64_t h = mix(0x1234, b.kind);
    if (b.kind == BK_SOLID) h = mix(h, b.value);
    if (b.kind == BK_PARTED) {
        const uint32_t* d = brick_data(b.value);
        for (uint32_t i = 0; i < vol_; ++i) h = mix(h, d[i]);
    }
    return h;
}

uint64_t HostOctree::hash_node(size_t key) const {
    const NodeRec& n = nodes_[key];
    uint64_t h = mix(0xABCD, n.kind);
    h = mix(h, stored_occupied_bits(key));
    if (n.kind == NK_INTERNAL) {
        for (int o = 0; o < 8; ++o) {
            const uint32_t c = n.link == LK_CHILDREN ? n.child[o] : NIL;
            h = mix(h, key_is_valid(c) ? hash_node(c) : 0x5EED);
        }
    } else if (n.kind == NK_LEAF) {
        for (int o = 0; o < 8; ++o) h = mix(h, hash_brick(n.brick[o]));
    } else if (n.kind == NK_UNIFORM) {
        h = mix(h, hash_brick(n.brick[0]));
    }
    return h;
}

uint64_t HostOctree::structure_hash() const {
    uint64_t h = mix(size_, dim_);
    for (const svx_albedo& a : colors_)
        h = mix(h, ((uint64_t)a.r << 24) | ((uint64_t)a.g << 16) | ((uint64_t)a.b << 8) | a.a);
    h = mix(h, 0xDA7A);
    for (uint32_t d : datas_) h = mix(h, d);
    return mix(h, hash_node(0));
}

}  // namespace svx
