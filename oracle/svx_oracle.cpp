// ORACLE — TEST INFRASTRUCTURE ONLY (see svx_oracle.hpp).
// Literal CPU restatement of the reference algorithms. Build with
//   g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math
// so that every f32 operation is a single IEEE-754 binary32 round-to-nearest
// operation in the reference's left-to-right order (Rust never contracts to FMA).
// All `file:line` citations are relative to /root/reference/.
#include "svx_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <unordered_map>

namespace svxo {

static constexpr uint32_t EMPTY_MARKER_U32 = 0xFFFFFFFFu;  // src/object_pool.rs:14-16
static constexpr uint32_t EMPTY_MARKER_U16 = 0xFFFFu;
static constexpr uint8_t OOB_OCTANT = 8;                   // src/spatial/lut.rs:154
static constexpr float FLOAT_ERROR_TOLERANCE = 0.00001f;   // src/spatial/raytracing/mod.rs:5
static constexpr size_t BITMAP_DIMENSION = 4;              // src/spatial/math/mod.rs:39

// ---------------------------------------------------------------------------------------------
// Rust scalar semantics
// ---------------------------------------------------------------------------------------------
// `f32 as i32`: saturating, NaN -> 0
static inline int32_t f2i32(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
// `f32 as usize`: saturating, NaN -> 0, negative -> 0
static inline size_t f2usize(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}
// `f32::signum`: 1.0 for +0.0 and positive, -1.0 for -0.0 and negative, NaN for NaN
static inline float signum(float v) {
    if (v != v) return v;
    return std::copysign(1.0f, v);
}
// `f32::clamp`
static inline float clampf(float v, float lo, float hi) {
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}
// `f32::min` / `f32::max`: NaN-ignoring (IEEE minNum / maxNum)
static inline float fmin_(float a, float b) { return std::fmin(a, b); }
static inline float fmax_(float a, float b) { return std::fmax(a, b); }

// V3c<f32> operators, src/spatial/math/vector.rs:195-256
static inline V3f operator+(V3f a, V3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3f operator-(V3f a, V3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3f operator*(V3f a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3f operator/(V3f a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline V3f unit(float s) { return {s, s, s}; }
// vector.rs:75-77
static inline float length(V3f v) { return std::sqrt((v.x * v.x) + (v.y * v.y) + (v.z * v.z)); }
// vector.rs:79-81
static inline V3f normalized(V3f v) { return v / length(v); }
// vector.rs:186-192
static inline V3f cross(V3f a, V3f b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// From<V3c<f32>> for V3c<usize>: round() then `as usize` (vector.rs:306-316)
struct V3s {
    size_t x, y, z;
};
static inline V3s to_usize(V3f v) { return {f2usize(std::round(v.x)), f2usize(std::round(v.y)), f2usize(std::round(v.z))}; }
static inline V3f to_f32(V3u v) { return {(float)v.x, (float)v.y, (float)v.z}; }
static inline V3f to_f32(V3s v) { return {(float)v.x, (float)v.y, (float)v.z}; }
// derived PartialOrd on V3c (vector.rs:3): lexicographic x, y, z
static inline bool lex_le(V3f a, V3f b) {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    return a.z <= b.z;
}

// ---------------------------------------------------------------------------------------------
// spatial/math
// ---------------------------------------------------------------------------------------------
uint8_t hash_region(V3f offset, float size_half) {
    return (uint8_t)((offset.x >= size_half) + (offset.z >= size_half) * 2 + (offset.y >= size_half) * 4);
}

uint8_t hash_direction(V3f direction) {
    V3f offset = unit(1.0f) + direction;
    return hash_region(offset, 1.0f);
}

size_t flat_projection(size_t x, size_t y, size_t z, size_t size) { return x + (y * size) + (z * size * size); }

size_t position_in_bitmap_64bits(size_t x, size_t y, size_t z, size_t brick_size) {
    return flat_projection(x * BITMAP_DIMENSION / brick_size, y * BITMAP_DIMENSION / brick_size,
                           z * BITMAP_DIMENSION / brick_size, BITMAP_DIMENSION);
}

void set_occupancy_in_bitmap_64bits(size_t px, size_t py, size_t pz, size_t size, size_t brick_dim, bool occupied,
                                    uint64_t* bitmap) {
    if (brick_dim == 1) {
        *bitmap = occupied ? UINT64_MAX : 0;
        return;
    }
    const size_t update_count = f2usize(std::ceil((float)size * (float)BITMAP_DIMENSION / (float)brick_dim));
    V3f startf = V3f{(float)(px * BITMAP_DIMENSION), (float)(py * BITMAP_DIMENSION), (float)(pz * BITMAP_DIMENSION)} /
                 (float)brick_dim;
    startf = {std::floor(startf.x), std::floor(startf.y), std::floor(startf.z)};
    const V3s start = to_usize(startf);
    for (size_t x = start.x; x < std::min(start.x + update_count, BITMAP_DIMENSION); ++x)
        for (size_t y = start.y; y < std::min(start.y + update_count, BITMAP_DIMENSION); ++y)
            for (size_t z = start.z; z < std::min(start.z + update_count, BITMAP_DIMENSION); ++z) {
                const uint64_t pos_mask = 1ull << position_in_bitmap_64bits(x, y, z, BITMAP_DIMENSION);
                if (occupied)
                    *bitmap |= pos_mask;
                else
                    *bitmap &= ~pos_mask;
            }
}

// src/spatial/math/mod.rs:44-77
static V3s matrix_index_for(const Cube& bounds, V3u position, uint32_t matrix_dimension) {
    V3f v = (to_f32(position) - bounds.min_position) * (float)matrix_dimension / bounds.size;
    v = {std::floor(v.x), std::floor(v.y), std::floor(v.z)};
    return to_usize(v);
}

// ---------------------------------------------------------------------------------------------
// LUTs: regenerated from the generator logic (src/spatial/lut.rs:12-152), never copied
// ---------------------------------------------------------------------------------------------
static Luts build_luts() {
    Luts l{};
    // OCTANT_OFFSET_REGION_LUT (lut.rs:156-197): inverse of hash_region (x -> bit0, z -> bit1, y -> bit2)
    for (int o = 0; o < 8; ++o) l.octant_offset[o] = {(float)(o & 1), (float)((o >> 2) & 1), (float)((o >> 1) & 1)};
    // BITMAP_MASK_FOR_OCTANT_LUT via convert_8bit_bitmap_to_64bit (lut.rs:12-37)
    for (int o = 0; o < 8; ++o) {
        size_t mx = (size_t)l.octant_offset[o].x * 2, my = (size_t)l.octant_offset[o].y * 2,
               mz = (size_t)l.octant_offset[o].z * 2;
        uint64_t bm = 0;
        for (size_t x = mx; x < mx + 2; ++x)
            for (size_t y = my; y < my + 2; ++y)
                for (size_t z = mz; z < mz + 2; ++z) set_occupancy_in_bitmap_64bits(x, y, z, 1, 4, true, &bm);
        l.bitmap_mask_for_octant[o] = bm;
    }
    // BITMAP_INDEX_LUT via generate_bitmap_flat_index_lut (lut.rs:139-150)
    for (size_t x = 0; x < 4; ++x)
        for (size_t y = 0; y < 4; ++y)
            for (size_t z = 0; z < 4; ++z) l.bitmap_index[x][y][z] = (uint8_t)position_in_bitmap_64bits(x, y, z, 4);
    // OCTANT_STEP_RESULT_LUT via generate_octant_step_result_lut (lut.rs:91-137)
    for (int octant = 0; octant < 8; ++octant) {
        const int shift = 4 * octant;
        for (int z = -1; z <= 1; ++z)
            for (int y = -1; y <= 1; ++y)
                for (int x = -1; x <= 1; ++x) {
                    const float SPACE = 12.0f;
                    V3f off = l.octant_offset[octant];
                    V3f center = {SPACE / 4.0f + off.x * (SPACE / 2.0f), SPACE / 4.0f + off.y * (SPACE / 2.0f),
                                  SPACE / 4.0f + off.z * (SPACE / 2.0f)};
                    V3f after = {center.x + (float)x * (SPACE / 2.0f), center.y + (float)y * (SPACE / 2.0f),
                                 center.z + (float)z * (SPACE / 2.0f)};
                    uint32_t res;
                    if (after.x < 0.0f || after.x > SPACE || after.y < 0.0f || after.y > SPACE || after.z < 0.0f ||
                        after.z > SPACE)
                        res = OOB_OCTANT;
                    else
                        res = hash_region(after, SPACE / 2.0f);
                    l.octant_step_result[x + 1][y + 1][z + 1] |= (res & 0x0Fu) << shift;
                }
    }
    // RAY_TO_NODE_OCCUPANCY_BITMASK_LUT via generate_lut_64_bits (lut.rs:39-89)
    for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y)
            for (int z = 0; z < 4; ++z) {
                const size_t pos = position_in_bitmap_64bits((size_t)x, (size_t)y, (size_t)z, 4);
                for (int dx = -1; dx <= 1; ++dx)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dz = -1; dz <= 1; ++dz) {
                            if (dx == 0 || dy == 0 || dz == 0) continue;
                            const uint8_t dir = hash_direction({(float)dx, (float)dy, (float)dz});
                            const int mvx = std::clamp(x + dx * 4, 0, 3), mvy = std::clamp(y + dy * 4, 0, 3),
                                      mvz = std::clamp(z + dz * 4, 0, 3);
                            uint64_t mask = 0;
                            for (int bx = std::min(mvx, x); bx <= std::max(mvx, x); ++bx)
                                for (int by = std::min(mvy, y); by <= std::max(mvy, y); ++by)
                                    for (int bz = std::min(mvz, z); bz <= std::max(mvz, z); ++bz)
                                        set_occupancy_in_bitmap_64bits((size_t)bx, (size_t)by, (size_t)bz, 1, 4, true,
                                                                       &mask);
                            l.ray_to_node_occupancy[pos][dir] = mask;
                        }
            }
    return l;
}

const Luts& luts() {
    static const Luts l = build_luts();
    return l;
}

// ---------------------------------------------------------------------------------------------
// spatial/mod.rs, spatial/raytracing/mod.rs
// ---------------------------------------------------------------------------------------------
Cube child_bounds_for(const Cube& c, uint8_t octant) {
    const float child_size = c.size / 2.0f;
    return Cube{c.min_position + (luts().octant_offset[octant] * child_size), child_size};
}

V3f cross_product(V3f a, V3f b) { return cross(a, b); }

// plane_line_intersection, src/spatial/raytracing/mod.rs:86-104. V3c::dot (vector.rs:182-184) sums left to right.
bool plane_line_intersection(V3f plane_point, V3f plane_normal, V3f line_origin, V3f line_direction, float* distance) {
    auto dot = [](V3f a, V3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
    const V3f origins_diff = plane_point - line_origin;
    const float plane_line_dot_to_plane = dot(origins_diff, plane_normal);
    const float directions_dot = dot(line_direction, plane_normal);
    if (0.0f == directions_dot) {
        // line and plane are parallel: distance 0 when the origin is already on the plane, otherwise no intersection
        if (0.0f == dot(origins_diff, plane_normal)) {
            *distance = 0.0f;
            return true;
        }
        return false;
    }
    *distance = plane_line_dot_to_plane / directions_dot;
    return true;
}

bool intersect_ray(const Cube& c, const Ray& ray, bool* has_distance, float* distance) {
    const V3f max_position = c.min_position + unit(c.size);
    const float t1 = (c.min_position.x - ray.origin.x) / ray.direction.x;
    const float t2 = (max_position.x - ray.origin.x) / ray.direction.x;
    const float t3 = (c.min_position.y - ray.origin.y) / ray.direction.y;
    const float t4 = (max_position.y - ray.origin.y) / ray.direction.y;
    const float t5 = (c.min_position.z - ray.origin.z) / ray.direction.z;
    const float t6 = (max_position.z - ray.origin.z) / ray.direction.z;

    const float tmin = fmax_(fmax_(fmin_(t1, t2), fmin_(t3, t4)), fmin_(t5, t6));
    const float tmax = fmin_(fmin_(fmax_(t1, t2), fmax_(t3, t4)), fmax_(t5, t6));

    if (tmax < 0.0f || tmin > tmax) return false;
    if (tmin < 0.0f) {
        *has_distance = false;
        *distance = 0.0f;
        return true;
    }
    *has_distance = true;
    *distance = tmin;
    return true;
}

uint8_t step_octant(uint8_t octant, V3f step) {
    auto sgn = [](int32_t v) { return (v > 0) - (v < 0); };
    const uint32_t shift = 4u * octant;
    const uint32_t v = luts().octant_step_result[sgn(f2i32(step.x)) + 1][sgn(f2i32(step.y)) + 1][sgn(f2i32(step.z)) + 1];
    return (uint8_t)((v & (0x0Fu << shift)) >> shift);
}

V3f cube_impact_normal(const Cube& c, V3f impact_point) {
    const V3f mid_to_impact = c.min_position + unit(c.size / 2.0f) - impact_point;
    const float max_component =
        fmax_(fmax_(std::fabs(mid_to_impact.x), std::fabs(mid_to_impact.y)), std::fabs(mid_to_impact.z));
    const V3f n = {std::fabs(mid_to_impact.x) == max_component ? -mid_to_impact.x : 0.0f,
                   std::fabs(mid_to_impact.y) == max_component ? -mid_to_impact.y : 0.0f,
                   std::fabs(mid_to_impact.z) == max_component ? -mid_to_impact.z : 0.0f};
    return normalized(n);
}

// ---------------------------------------------------------------------------------------------
// object_pool.rs
// ---------------------------------------------------------------------------------------------
bool NodePool::is_next_available() const {
    return first_available + 1 < item.size() && !reserved[first_available + 1];
}
bool NodePool::check_first_available() {
    if (first_available < item.size() && !reserved[first_available]) return true;
    if (is_next_available()) {
        first_available += 1;
        return true;
    }
    first_available = item.size();
    return false;
}
size_t NodePool::allocate() {
    size_t key;
    if (check_first_available()) {
        reserved[first_available] = 1;
        key = first_available;
    } else {
        item.emplace_back();
        reserved.push_back(1);
        key = item.size() - 1;
    }
    if (is_next_available()) first_available += 1;
    return key;
}
size_t NodePool::push(Node&& n) {
    const size_t key = allocate();
    item[key] = std::move(n);
    return key;
}
bool NodePool::free_key(size_t key) {
    if (!key_is_valid(key)) return false;
    reserved[key] = 0;
    first_available = std::min(first_available, key);
    return true;
}
void NodePool::swap_items(size_t a, size_t b) {
    std::swap(item[a], item[b]);
    uint8_t t = reserved[a];
    reserved[a] = reserved[b];
    reserved[b] = t;
}

// ---------------------------------------------------------------------------------------------
// octree/node.rs : palette index values + BrickData helpers
// ---------------------------------------------------------------------------------------------
bool Brick::operator==(const Brick& o) const {
    if (kind != o.kind) return false;
    if (kind == BrickKind::Solid) return solid == o.solid;
    if (kind == BrickKind::Parted) return data == o.data;
    return true;
}

static inline uint32_t pix_visual(uint16_t c) { return (uint32_t)c | (EMPTY_MARKER_U16 << 16); }       // node.rs:354
static inline uint32_t pix_informal(uint16_t d) { return EMPTY_MARKER_U16 | ((uint32_t)d << 16); }     // node.rs:358
static inline uint32_t pix_complex(uint16_t c, uint16_t d) { return (uint32_t)c | ((uint32_t)d << 16); }  // :362
static inline size_t pix_color_index(uint32_t i) { return i & 0x0000FFFFu; }
static inline size_t pix_data_index(uint32_t i) { return (i & 0xFFFF0000u) >> 16; }
static inline bool pix_color_is_some(uint32_t i) { return pix_color_index(i) < EMPTY_MARKER_U16; }     // node.rs:389
static inline bool pix_color_is_none(uint32_t i) { return !pix_color_is_some(i); }
static inline bool pix_data_is_none(uint32_t i) { return pix_data_index(i) == EMPTY_MARKER_U16; }      // node.rs:397
static inline bool pix_data_is_some(uint32_t i) { return !pix_data_is_none(i); }
static inline uint32_t pix_overwrite_color(uint32_t i, uint32_t d) { return (i & 0xFFFF0000u) | (d & 0x0000FFFFu); }
static inline uint32_t pix_overwrite_data(uint32_t i, uint32_t d) { return (i & 0x0000FFFFu) | (d & 0xFFFF0000u); }

// node.rs:405-427 ; VoxelData for u32: is_empty <=> == 0 (detail.rs:40-44)
bool Octree::pix_points_to_empty(uint32_t index) const {
    return (pix_color_is_none(index) || voxel_color_palette[pix_color_index(index)].a == 0) &&
           (pix_data_is_none(index) || voxel_data_palette[pix_data_index(index)] == 0);
}

// node.rs:429-467
Entry Octree::pix_get_ref(uint32_t index) const {
    Entry e;
    if (pix_data_is_none(index) && pix_color_is_none(index)) return e;
    if (pix_data_is_none(index)) {
        e.kind = EntryKind::Visual;
        e.albedo = voxel_color_palette[pix_color_index(index)];
        return e;
    }
    if (pix_color_is_none(index)) {
        e.kind = EntryKind::Informative;
        e.data = voxel_data_palette[pix_data_index(index)];
        return e;
    }
    e.kind = EntryKind::Complex;
    e.albedo = voxel_color_palette[pix_color_index(index)];
    e.data = voxel_data_palette[pix_data_index(index)];
    return e;
}

// node.rs:244-272
uint64_t Octree::brick_occupied_bits(const std::vector<uint32_t>& brick) const {
    uint64_t bitmap = 0;
    const size_t d = brick_dim;
    for (size_t x = 0; x < d; ++x)
        for (size_t y = 0; y < d; ++y)
            for (size_t z = 0; z < d; ++z)
                if (!pix_points_to_empty(brick[flat_projection(x, y, z, d)]))
                    set_occupancy_in_bitmap_64bits(x, y, z, 1, d, true, &bitmap);
    return bitmap;
}

// node.rs:275-297
uint64_t Octree::calculate_occupied_bits(const Brick& b) const {
    switch (b.kind) {
        case BrickKind::Empty: return 0;
        case BrickKind::Solid: return pix_points_to_empty(b.solid) ? 0 : UINT64_MAX;
        case BrickKind::Parted: return brick_occupied_bits(b.data);
    }
    return 0;
}

// node.rs:300-313
static const uint32_t* get_homogeneous_data(const Brick& b) {
    switch (b.kind) {
        case BrickKind::Empty: return nullptr;
        case BrickKind::Solid: return &b.solid;
        case BrickKind::Parted:
            // same answer as the reference's linear scan; `witness` remembers an index that differed last time so
            // that repeated calls on a growing brick stay O(1) (build-time only, no effect on results)
            if (b.witness != 0 && b.witness < b.data.size() && b.data[b.witness] != b.data[0]) return nullptr;
            for (size_t i = 1; i < b.data.size(); ++i)
                if (b.data[i] != b.data[0]) {
                    b.witness = (uint32_t)i;
                    return nullptr;
                }
            b.witness = 0;
            return &b.data[0];
    }
    return nullptr;
}

// node.rs:316-331
bool Octree::brick_simplify(Brick& b) const {
    const uint32_t* h = get_homogeneous_data(b);
    if (!h) return false;
    const uint32_t v = *h;
    if (pix_points_to_empty(v)) {
        b.kind = BrickKind::Empty;
        b.data = std::vector<uint32_t>();
    } else {
        b.kind = BrickKind::Solid;
        b.solid = v;
        b.data = std::vector<uint32_t>();
    }
    return true;
}

// node.rs:518-552
bool Octree::node_is_all(const Node& n, uint32_t data) const {
    auto brick_all = [&](const Brick& b) {
        switch (b.kind) {
            case BrickKind::Empty: return false;
            case BrickKind::Solid: return b.solid == data;
            case BrickKind::Parted: {
                const uint32_t* h = get_homogeneous_data(b);
                return h ? (*h == data) : false;
            }
        }
        return false;
    };
    switch (n.kind) {
        case NodeKind::UniformLeaf: return brick_all(n.ubrick);
        case NodeKind::Leaf:
            for (int o = 0; o < 8; ++o)
                if (!brick_all(n.bricks[o])) return false;
            return true;
        default: return false;
    }
}

// node.rs:554-573
static bool node_compare(const Node& a, const Node& b) {
    switch (a.kind) {
        case NodeKind::Nothing: return b.kind == NodeKind::Nothing;
        case NodeKind::Internal: return false;
        case NodeKind::UniformLeaf: return b.kind == NodeKind::UniformLeaf && a.ubrick == b.ubrick;
        case NodeKind::Leaf:
            if (b.kind != NodeKind::Leaf) return false;
            for (int o = 0; o < 8; ++o)
                if (!(a.bricks[o] == b.bricks[o])) return false;
            return true;
    }
    return false;
}

// node.rs:49-54
static inline size_t child_of(const Children& c, uint8_t octant) {
    if (c.kind == ChildrenKind::Children) return c.child[octant];
    return (size_t)-1;  // empty_marker::<usize>()
}
// node.rs:56-64
static inline uint32_t* child_mut(Children& c, size_t index) {
    if (c.kind == ChildrenKind::NoChildren) {
        c.kind = ChildrenKind::Children;
        for (int i = 0; i < 8; ++i) c.child[i] = EMPTY_MARKER_U32;
    }
    // the reference panics when this is an OccupancyBitmap
    return &c.child[index];
}

// ---------------------------------------------------------------------------------------------
// octree/mod.rs
// ---------------------------------------------------------------------------------------------
// mod.rs:173-205
Status Octree::create(uint32_t size, uint32_t brick_dimension, Octree** out) {
    *out = nullptr;
    auto fract_nonzero = [](float v) {
        // f32::fract() != 0.0 ; log2 of 0 is -inf whose fract is NaN (!= 0.0 is true)
        float ip;
        float fr = std::modf(v, &ip);
        if (std::isinf(v)) fr = NAN;
        return fr != 0.0f;
    };
    // (x as f32).log(2.0) is ln(x)/ln(2) in Rust's std; powers of two below 2^24 evaluate to exact integers
    auto log_2 = [](float v) { return std::log(v) / std::log(2.0f); };
    if (0 == size || fract_nonzero(log_2((float)brick_dimension))) return E_INVALID_BRICK_DIMENSION;
    if (brick_dimension > size || 0 == size || fract_nonzero(log_2((float)size / (float)brick_dimension)))
        return E_INVALID_SIZE;
    if (size < brick_dimension * 2) return E_INVALID_STRUCTURE;
    Octree* t = new Octree();
    t->octree_size = size;
    t->brick_dim = brick_dimension;
    t->auto_simplify = true;
    const size_t root = t->nodes.push(Node());
    (void)root;  // == 0
    t->node_children.assign(1, Children());
    *out = t;
    return OK;
}

// detail.rs:20-27
static bool bound_contains(const Cube& b, V3f p) {
    return p.x >= b.min_position.x && p.x < b.min_position.x + b.size && p.y >= b.min_position.y &&
           p.y < b.min_position.y + b.size && p.z >= b.min_position.z && p.z < b.min_position.z + b.size;
}
// detail.rs:30-38
static uint8_t child_octant_for(const Cube& b, V3f p) { return hash_region(p - b.min_position, b.size / 2.0f); }

// mod.rs:209-371
Entry Octree::get(V3u position_u) const { return get_internal(0, Cube{unit(0.0f), (float)octree_size}, position_u); }

// mod.rs:220-371
Entry Octree::get_internal(size_t current_node_key, Cube current_bounds, V3u position_u) const {
    const V3f position = to_f32(position_u);
    if (!bound_contains(current_bounds, position)) return Entry();
    for (;;) {
        const Node& node = nodes.item[current_node_key];
        switch (node.kind) {
            case NodeKind::Nothing: return Entry();
            case NodeKind::Leaf: {
                const uint8_t oct = child_octant_for(current_bounds, position);
                const Brick& b = node.bricks[oct];
                switch (b.kind) {
                    case BrickKind::Empty: return Entry();
                    case BrickKind::Parted: {
                        current_bounds = child_bounds_for(current_bounds, oct);
                        // V3c::from(position): f32 -> u32 by round (vector.rs:326-336)
                        const V3s mi = matrix_index_for(current_bounds, position_u, brick_dim);
                        const size_t fi = flat_projection(mi.x, mi.y, mi.z, brick_dim);
                        if (!pix_points_to_empty(b.data[fi])) return pix_get_ref(b.data[fi]);
                        return Entry();
                    }
                    case BrickKind::Solid: return pix_get_ref(b.solid);
                }
                return Entry();
            }
            case NodeKind::UniformLeaf: {
                const Brick& b = node.ubrick;
                switch (b.kind) {
                    case BrickKind::Empty: return Entry();
                    case BrickKind::Parted: {
                        const V3s mi = matrix_index_for(current_bounds, position_u, brick_dim);
                        const size_t fi = flat_projection(mi.x, mi.y, mi.z, brick_dim);
                        if (pix_points_to_empty(b.data[fi])) return Entry();
                        return pix_get_ref(b.data[fi]);
                    }
                    case BrickKind::Solid:
                        if (pix_points_to_empty(b.solid)) return Entry();
                        return pix_get_ref(b.solid);
                }
                return Entry();
            }
            case NodeKind::Internal: {
                const uint8_t oct = child_octant_for(current_bounds, position);
                const size_t child = child_of(node_children[current_node_key], oct);
                if (nodes.key_is_valid(child)) {
                    current_node_key = child;
                    current_bounds = child_bounds_for(current_bounds, oct);
                } else {
                    return Entry();
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// octree/detail.rs
// ---------------------------------------------------------------------------------------------
// detail.rs:524-544
uint64_t Octree::stored_occupied_bits(size_t node_key) const {
    const Node& n = nodes.item[node_key];
    switch (n.kind) {
        case NodeKind::Leaf:
        case NodeKind::UniformLeaf: {
            const Children& c = node_children[node_key];
            if (c.kind == ChildrenKind::OccupancyBitmap) return c.bitmap;
            return 0;
        }
        case NodeKind::Nothing: return 0;
        case NodeKind::Internal: return n.occupied_bits;
    }
    return 0;
}

// detail.rs:547-569
void Octree::store_occupied_bits(size_t node_key, uint64_t bits) {
    Node& n = nodes.item[node_key];
    Children& c = node_children[node_key];
    switch (n.kind) {
        case NodeKind::Internal: n.occupied_bits = bits; break;
        case NodeKind::Nothing:
            c.kind = ChildrenKind::OccupancyBitmap;
            c.bitmap = bits;
            break;
        case NodeKind::Leaf:
        case NodeKind::UniformLeaf:
            // NoChildren / OccupancyBitmap -> OccupancyBitmap(bits); Children(_) panics in the reference
            c.kind = ChildrenKind::OccupancyBitmap;
            c.bitmap = bits;
            break;
    }
}

// detail.rs:489-500
Brick Octree::try_brick_from_node(size_t node_key) const {
    if (!nodes.key_is_valid(node_key)) return Brick();
    const Node& n = nodes.item[node_key];
    if (n.kind == NodeKind::UniformLeaf) return n.ubrick;
    return Brick();
}

// detail.rs:503-520
void Octree::deallocate_children_of(size_t node) {
    if (!nodes.key_is_valid(node)) return;
    if (node_children[node].kind != ChildrenKind::Children) return;
    std::vector<size_t> to_deallocate;
    for (int i = 0; i < 8; ++i) {
        const size_t child = node_children[node].child[i];
        if (nodes.key_is_valid(child)) to_deallocate.push_back(child);
    }
    for (size_t child : to_deallocate) {
        deallocate_children_of(child);
        nodes.free_key(child);
        node_children[child] = Children();
    }
}

// update/mod.rs:563-628
static void dilute_brick_data(const std::vector<uint32_t>& brick_data, uint32_t brick_dim,
                              std::vector<uint32_t> result[8]) {
    const size_t n = (size_t)brick_dim * brick_dim * brick_dim;
    if (1 == brick_dim) {
        for (int o = 0; o < 8; ++o) result[o] = brick_data;
        return;
    }
    for (int o = 0; o < 8; ++o) result[o].assign(n, brick_data[o]);
    if (2 == brick_dim) return;
    const Luts& l = luts();
    for (size_t octant = 0; octant < 8; ++octant) {
        const V3s off = to_usize(l.octant_offset[octant]);
        const V3s brick_offset = {off.x * 2, off.y * 2, off.z * 2};
        const size_t first = flat_projection(brick_offset.x, brick_offset.y, brick_offset.z, brick_dim);
        std::vector<uint32_t> nb(n, brick_data[first]);
        for (size_t x = 0; x < brick_dim; ++x)
            for (size_t y = 0; y < brick_dim; ++y)
                for (size_t z = 0; z < brick_dim; ++z) {
                    if (x < 2 && y < 2 && z < 2) continue;
                    nb[flat_projection(x, y, z, brick_dim)] =
                        brick_data[flat_projection(brick_offset.x + x / 2, brick_offset.y + y / 2,
                                                   brick_offset.z + z / 2, brick_dim)];
                }
        result[octant] = std::move(nb);
    }
}

// detail.rs:321-486
void Octree::subdivide_leaf_to_nodes(size_t node_key, size_t target_octant) {
    // panics in the reference unless node_children[node_key] is an OccupancyBitmap
    Node node_content;
    node_content.kind = NodeKind::Internal;
    node_content.occupied_bits = node_children[node_key].bitmap;
    std::swap(node_content, nodes.item[node_key]);
    uint32_t node_new_children[8];
    for (int i = 0; i < 8; ++i) node_new_children[i] = EMPTY_MARKER_U32;

    auto grow_children = [&](uint32_t key) {
        if (node_children.size() < (size_t)key + 1) node_children.resize((size_t)key + 1);
    };

    if (node_content.kind == NodeKind::Leaf) {
        for (size_t octant = 0; octant < 8; ++octant) {
            Brick brick;
            std::swap(brick, node_content.bricks[octant]);
            switch (brick.kind) {
                case BrickKind::Empty:
                    if (octant == target_octant) {
                        node_new_children[octant] = (uint32_t)nodes.push(Node());
                        grow_children(node_new_children[octant]);
                    }
                    break;
                case BrickKind::Solid: {
                    Node n;
                    n.kind = NodeKind::UniformLeaf;
                    n.ubrick = brick;
                    node_new_children[octant] = (uint32_t)nodes.push(std::move(n));
                    grow_children(node_new_children[octant]);
                    Children& c = node_children[node_new_children[octant]];
                    c.kind = ChildrenKind::OccupancyBitmap;
                    c.bitmap = UINT64_MAX;
                    break;
                }
                case BrickKind::Parted: {
                    Node n;
                    n.kind = NodeKind::UniformLeaf;
                    n.ubrick = brick;  // brick.clone()
                    node_new_children[octant] = (uint32_t)nodes.push(std::move(n));
                    grow_children(node_new_children[octant]);
                    Children& c = node_children[node_new_children[octant]];
                    c.kind = ChildrenKind::OccupancyBitmap;
                    // detail.rs:405: computed from `bricks[octant]`, which was swapped to Empty at :343 -> 0
                    c.bitmap = calculate_occupied_bits(node_content.bricks[octant]);
                    break;
                }
            }
        }
    } else if (node_content.kind == NodeKind::UniformLeaf) {
        Brick& brick = node_content.ubrick;
        switch (brick.kind) {
            case BrickKind::Empty: {
                node_new_children[target_octant] = (uint32_t)nodes.push(Node());
                grow_children(node_new_children[target_octant]);
                Children& c = node_children[node_new_children[target_octant]];
                c.kind = ChildrenKind::OccupancyBitmap;
                c.bitmap = 0;
                break;
            }
            case BrickKind::Solid:
                for (size_t octant = 0; octant < 8; ++octant) {
                    Node n;
                    n.kind = NodeKind::UniformLeaf;
                    n.ubrick.kind = BrickKind::Solid;
                    n.ubrick.solid = brick.solid;
                    node_new_children[octant] = (uint32_t)nodes.push(std::move(n));
                    grow_children(node_new_children[octant]);
                    Children& c = node_children[node_new_children[octant]];
                    c.kind = ChildrenKind::OccupancyBitmap;
                    c.bitmap = UINT64_MAX;
                }
                break;
            case BrickKind::Parted: {
                std::vector<uint32_t> children_bricks[8];
                dilute_brick_data(brick.data, brick_dim, children_bricks);
                for (size_t octant = 0; octant < 8; ++octant) {
                    const uint64_t child_bits = brick_occupied_bits(children_bricks[octant]);
                    Node n;
                    n.kind = NodeKind::UniformLeaf;
                    n.ubrick.kind = BrickKind::Parted;
                    n.ubrick.data = std::move(children_bricks[octant]);
                    node_new_children[octant] = (uint32_t)nodes.push(std::move(n));
                    grow_children(node_new_children[octant]);
                    Children& c = node_children[node_new_children[octant]];
                    c.kind = ChildrenKind::OccupancyBitmap;
                    c.bitmap = child_bits;
                }
                break;
            }
        }
    }
    // Nothing | Internal: the reference panics ("Non-leaf node expected to be Leaf")
    Children& c = node_children[node_key];
    c.kind = ChildrenKind::Children;
    for (int i = 0; i < 8; ++i) c.child[i] = node_new_children[i];
}

// ---------------------------------------------------------------------------------------------
// octree/update/mod.rs
// ---------------------------------------------------------------------------------------------
// update/mod.rs:55-136
uint32_t Octree::add_to_palette(const Entry& e) {
    auto albedo_key = [](Albedo a) { return ((uint32_t)a.r << 24) | ((uint32_t)a.g << 16) | ((uint32_t)a.b << 8) | a.a; };
    auto albedo_zero = [](Albedo a) { return a.r == 0 && a.g == 0 && a.b == 0 && a.a == 0; };
    auto color_index = [&](Albedo a) -> size_t {
        const uint32_t k = albedo_key(a);
        auto it = color_lookup_.find(k);
        if (it != color_lookup_.end()) return it->second;
        const size_t idx = voxel_color_palette.size();
        color_lookup_.emplace(k, idx);
        voxel_color_palette.push_back(a);
        return idx;
    };
    auto data_index = [&](uint32_t d) -> size_t {
        auto it = data_lookup_.find(d);
        if (it != data_lookup_.end()) return it->second;
        const size_t idx = voxel_data_palette.size();
        data_lookup_.emplace(d, idx);
        voxel_data_palette.push_back(d);
        return idx;
    };
    switch (e.kind) {
        case EntryKind::Empty: return EMPTY_MARKER_U32;
        case EntryKind::Visual:
            if (albedo_zero(e.albedo)) return EMPTY_MARKER_U32;
            return pix_visual((uint16_t)color_index(e.albedo));
        case EntryKind::Informative:
            if (e.data == 0) return EMPTY_MARKER_U32;
            return pix_informal((uint16_t)data_index(e.data));
        case EntryKind::Complex: {
            if (albedo_zero(e.albedo)) {
                Entry i;
                i.kind = EntryKind::Informative;
                i.data = e.data;
                return add_to_palette(i);
            } else if (e.data == 0) {
                Entry v;
                v.kind = EntryKind::Visual;
                v.albedo = e.albedo;
                return add_to_palette(v);
            }
            const size_t ci = color_index(e.albedo);
            const size_t di = data_index(e.data);
            return pix_complex((uint16_t)ci, (uint16_t)di);
        }
    }
    return EMPTY_MARKER_U32;
}

// update/mod.rs:637-675
static size_t update_brick(bool overwrite_if_empty, std::vector<uint32_t>& brick, const Cube& brick_bounds,
                           uint32_t brick_dim, V3u position, uint32_t size, uint32_t data) {
    const V3s mi = matrix_index_for(brick_bounds, position, brick_dim);
    const size_t update_size = std::min((size_t)brick_dim - mi.x, (size_t)size);
    for (size_t x = mi.x; x < std::min(mi.x + size, (size_t)brick_dim); ++x)
        for (size_t y = mi.y; y < std::min(mi.y + size, (size_t)brick_dim); ++y)
            for (size_t z = mi.z; z < std::min(mi.z + size, (size_t)brick_dim); ++z) {
                const size_t fi = flat_projection(x, y, z, brick_dim);
                if (overwrite_if_empty) {
                    brick[fi] = data;
                } else {
                    if (pix_color_is_some(data)) brick[fi] = pix_overwrite_color(brick[fi], data);
                    if (pix_data_is_some(data)) brick[fi] = pix_overwrite_data(brick[fi], data);
                }
            }
    return update_size;
}

// ---- the block cache of Voxels (svx_oracle.hpp) ----
static bool block_is_mixed(const std::vector<uint32_t>& v, size_t bx, size_t by, size_t bz, size_t d) {
    const uint32_t v0 = v[flat_projection(bx * 2, by * 2, bz * 2, d)];
    for (size_t c = 1; c < 8; ++c)
        if (v[flat_projection(bx * 2 + (c & 1), by * 2 + ((c >> 1) & 1), bz * 2 + (c >> 2), d)] != v0) return true;
    return false;
}
static void rebuild_blocks(const Voxels& v, size_t d) {
    const size_t half = d / 2, blocks = half * half * half;
    v.mixed_blocks.assign((blocks + 63) / 64 + 1, 0ull);
    uint32_t count = 0;
    for (size_t bz = 0; bz < half; ++bz)
        for (size_t by = 0; by < half; ++by)
            for (size_t bx = 0; bx < half; ++bx)
                if (block_is_mixed(v, bx, by, bz, d)) {
                    const size_t i = (bz * half + by) * half + bx;
                    v.mixed_blocks[i >> 6] |= 1ull << (i & 63);
                    ++count;
                }
    v.mixed_count = count;
}
// An in-place edit of a brick that lives in the tree: the literal update above, then the blocks the written box touches are
// looked at again. (Overload resolution picks this one for every Voxels argument, so a tree brick cannot be edited past it.)
static size_t update_brick(bool overwrite_if_empty, Voxels& brick, const Cube& brick_bounds, uint32_t brick_dim, V3u position,
                           uint32_t size, uint32_t data) {
    const size_t us = update_brick(overwrite_if_empty, static_cast<std::vector<uint32_t>&>(brick), brick_bounds, brick_dim, position,
                                   size, data);
    const size_t d = brick_dim, half = d / 2;
    if (brick.mixed_count == Voxels::UNKNOWN || half == 0) return us;
    const V3s mi = matrix_index_for(brick_bounds, position, brick_dim);
    const size_t hx = std::min(mi.x + size, d), hy = std::min(mi.y + size, d), hz = std::min(mi.z + size, d);
    if (mi.x >= hx || mi.y >= hy || mi.z >= hz) return us;
    for (size_t bz = mi.z / 2; bz <= (hz - 1) / 2; ++bz)
        for (size_t by = mi.y / 2; by <= (hy - 1) / 2; ++by)
            for (size_t bx = mi.x / 2; bx <= (hx - 1) / 2; ++bx) {
                const size_t i = (bz * half + by) * half + bx;
                const uint64_t bit = 1ull << (i & 63);
                const bool was = (brick.mixed_blocks[i >> 6] & bit) != 0, now = block_is_mixed(brick, bx, by, bz, d);
                if (was != now) {
                    brick.mixed_blocks[i >> 6] ^= bit;
                    brick.mixed_count += now ? 1u : 0xFFFFFFFFu;
                }
            }
    return us;
}

// update/mod.rs:160-549
size_t Octree::leaf_update(bool overwrite_if_empty, size_t node_key, const Cube& node_bounds, const Cube& target_bounds,
                           size_t target_child_octant, V3u position, uint32_t size, uint32_t target_content) {
    const size_t vol = (size_t)brick_dim * brick_dim * brick_dim;
    Node& node = nodes.item[node_key];
    switch (node.kind) {
        case NodeKind::Leaf: {
            Brick& b = node.bricks[target_child_octant];
            switch (b.kind) {
                case BrickKind::Empty: {
                    std::vector<uint32_t> new_brick(vol, EMPTY_MARKER_U32);
                    const size_t us =
                        update_brick(overwrite_if_empty, new_brick, target_bounds, brick_dim, position, size, target_content);
                    b.kind = BrickKind::Parted;
                    b.data = std::move(new_brick);
                    return us;
                }
                case BrickKind::Solid: {
                    const uint32_t voxel = b.solid;
                    size_t us;
                    if ((pix_points_to_empty(target_content) && !pix_points_to_empty(voxel)) ||
                        (!pix_points_to_empty(target_content) && voxel != target_content)) {
                        std::vector<uint32_t> new_brick(vol, voxel);
                        us = update_brick(overwrite_if_empty, new_brick, target_bounds, brick_dim, position, size,
                                          target_content);
                        b.kind = BrickKind::Parted;
                        b.data = std::move(new_brick);
                    } else {
                        us = 0;
                    }
                    return us;
                }
                case BrickKind::Parted:
                    return update_brick(overwrite_if_empty, b.data, target_bounds, brick_dim, position, size,
                                        target_content);
            }
            return 0;
        }
        case NodeKind::UniformLeaf: {
            Brick& mat = node.ubrick;
            switch (mat.kind) {
                case BrickKind::Empty: {
                    if (!pix_points_to_empty(target_content)) {
                        std::vector<uint32_t> new_brick(vol, EMPTY_MARKER_U32);  // add_to_palette(Empty) == empty marker
                        const size_t us = update_brick(overwrite_if_empty, new_brick, target_bounds, brick_dim, position,
                                                       size, target_content);
                        Node leaf;
                        leaf.kind = NodeKind::Leaf;
                        leaf.bricks[target_child_octant].kind = BrickKind::Parted;
                        leaf.bricks[target_child_octant].data = std::move(new_brick);
                        nodes.item[node_key] = std::move(leaf);
                        return us;
                    }
                    break;  // falls to the recursive call at update/mod.rs:486 (unbounded in the reference)
                }
                case BrickKind::Solid: {
                    const uint32_t voxel = mat.solid;
                    if (pix_points_to_empty(target_content) && pix_points_to_empty(voxel)) {
                        nodes.item[node_key] = Node();
                        return 0;
                    }
                    if ((!pix_points_to_empty(target_content) && voxel != target_content) ||
                        (pix_points_to_empty(target_content) && !pix_points_to_empty(voxel))) {
                        mat.kind = BrickKind::Parted;
                        mat.data.assign(vol, voxel);
                        return leaf_update(overwrite_if_empty, node_key, node_bounds, target_bounds, target_child_octant,
                                           position, size, target_content);
                    }
                    return 0;
                }
                case BrickKind::Parted: {
                    const V3s mi3 = matrix_index_for(node_bounds, position, brick_dim);
                    const size_t mi = flat_projection(mi3.x, mi3.y, mi3.z, brick_dim);
                    if (1 < brick_dim &&
                        ((pix_points_to_empty(target_content) && pix_points_to_empty(mat.data[mi])) ||
                         (!pix_points_to_empty(target_content) && mat.data[mi] == target_content))) {
                        return 0;
                    }
                    if (node_bounds.size <= (float)brick_dim && brick_dim > 1) {
                        return update_brick(overwrite_if_empty, mat.data, node_bounds, brick_dim, position, size,
                                            target_content);
                    }
                    Node leaf;
                    leaf.kind = NodeKind::Leaf;
                    size_t us = 0;
                    if (1 == brick_dim) {
                        for (int o = 0; o < 8; ++o) {
                            leaf.bricks[o].kind = BrickKind::Parted;
                            leaf.bricks[o].data = mat.data;
                        }
                        std::vector<uint32_t> new_brick = mat.data;
                        us = update_brick(overwrite_if_empty, new_brick, target_bounds, brick_dim, position, size,
                                          target_content);
                        leaf.bricks[target_child_octant].data = std::move(new_brick);
                    } else {
                        std::vector<uint32_t> child_bricks[8];
                        dilute_brick_data(mat.data, brick_dim, child_bricks);
                        for (size_t octant = 0; octant < 8; ++octant) {
                            if (octant == target_child_octant) {
                                us = update_brick(overwrite_if_empty, child_bricks[octant], target_bounds, brick_dim,
                                                  position, size, target_content);
                            }
                            leaf.bricks[octant].kind = BrickKind::Parted;
                            leaf.bricks[octant].data = std::move(child_bricks[octant]);
                        }
                    }
                    nodes.item[node_key] = std::move(leaf);
                    return us;
                }
            }
            // update/mod.rs:486-495: only reachable from UniformLeaf(Empty) with an empty target, where the
            // reference recurses without changing state (stack overflow). insert() never gets here because empty
            // entries return early (insert.rs:117-119).
            return 0;
        }
        case NodeKind::Internal: {
            // update/mod.rs:497-521 ("might induce data loss - see #69")
            Children& c = node_children[node_key];
            c.kind = ChildrenKind::OccupancyBitmap;
            c.bitmap = node.occupied_bits;
            Node leaf;
            leaf.kind = NodeKind::Leaf;
            for (uint8_t o = 0; o < 8; ++o) leaf.bricks[o] = try_brick_from_node(child_of(node_children[node_key], o));
            nodes.item[node_key] = std::move(leaf);
            deallocate_children_of(node_key);
            return leaf_update(overwrite_if_empty, node_key, node_bounds, target_bounds, target_child_octant, position,
                               size, target_content);
        }
        case NodeKind::Nothing: {
            Node leaf;
            leaf.kind = NodeKind::Leaf;
            for (uint8_t o = 0; o < 8; ++o) leaf.bricks[o] = try_brick_from_node(child_of(node_children[node_key], o));
            nodes.item[node_key] = std::move(leaf);
            deallocate_children_of(node_key);
            return leaf_update(overwrite_if_empty, node_key, node_bounds, target_bounds, target_child_octant, position,
                               size, target_content);
        }
    }
    return 0;
}

// update/mod.rs:689-1068
bool Octree::simplify(size_t node_key) {
    if (!nodes.key_is_valid(node_key)) return false;
    Node& node = nodes.item[node_key];
    const size_t d = brick_dim;
    switch (node.kind) {
        case NodeKind::Nothing: return true;
        case NodeKind::UniformLeaf: {
            Brick& brick = node.ubrick;
            switch (brick.kind) {
                case BrickKind::Empty: return true;
                case BrickKind::Solid:
                    if (pix_points_to_empty(brick.solid)) {
                        nodes.item[node_key] = Node();
                        node_children[node_key] = Children();
                        return true;
                    }
                    return false;
                case BrickKind::Parted: return brick_simplify(brick);
            }
            return false;
        }
        case NodeKind::Leaf: {
            bool simplified = false;
            bool is_leaf_uniform_solid = true;
            bool have_uniform_value = false;
            uint32_t uniform_solid_value = 0;
            for (int octant = 0; octant < 8; ++octant) {
                simplified |= brick_simplify(node.bricks[octant]);
                if (is_leaf_uniform_solid) {
                    if (node.bricks[octant].kind == BrickKind::Solid) {
                        if (have_uniform_value) {
                            if (uniform_solid_value != node.bricks[octant].solid) is_leaf_uniform_solid = false;
                        } else {
                            have_uniform_value = true;
                            uniform_solid_value = node.bricks[octant].solid;
                        }
                    } else {
                        is_leaf_uniform_solid = false;
                    }
                }
            }
            if (is_leaf_uniform_solid) {
                Node u;
                u.kind = NodeKind::UniformLeaf;
                u.ubrick.kind = BrickKind::Solid;
                u.ubrick.solid = uniform_solid_value;
                nodes.item[node_key] = std::move(u);
                return true;
            }
            // update/mod.rs:860-997. The reference fills the unified brick while it checks; the check is done first here
            // (same predicate, same early exits) so that the dim^3 buffer is only built when it is going to be used.
            bool is_leaf_uniform = true;
            const Luts& l = luts();
            const size_t brick_half = d / 2;
            for (int octant = 0; octant < 8 && is_leaf_uniform; ++octant) {
                const Brick& b = node.bricks[octant];
                if (b.kind != BrickKind::Parted) {
                    is_leaf_uniform &= (b == node.bricks[0]);
                    continue;
                }
                // "every aligned 2x2x2 block of the brick is one value" (the reference scans the brick, :884-980), from the
                // brick's block cache
                if (brick_half > 0) {
                    if (b.data.mixed_count == Voxels::UNKNOWN) rebuild_blocks(b.data, d);
                    if (b.data.mixed_count != 0) is_leaf_uniform = false;
                }
            }
            if (is_leaf_uniform) {
                std::vector<uint32_t> unified(d * d * d, EMPTY_MARKER_U32);
                for (int octant = 0; octant < 8; ++octant) {
                    const V3s octant_offset = to_usize(l.octant_offset[octant] * (float)brick_half);
                    const Brick& b = node.bricks[octant];
                    if (b.kind == BrickKind::Empty) continue;
                    for (size_t x = 0; x < brick_half; ++x)
                        for (size_t y = 0; y < brick_half; ++y)
                            for (size_t z = 0; z < brick_half; ++z)
                                unified[flat_projection(octant_offset.x + x, octant_offset.y + y, octant_offset.z + z, d)] =
                                    b.kind == BrickKind::Solid ? b.solid : b.data[flat_projection(x * 2, y * 2, z * 2, d)];
                }
                Node u;
                u.kind = NodeKind::UniformLeaf;
                u.ubrick.kind = BrickKind::Parted;
                u.ubrick.data = std::move(unified);
                nodes.item[node_key] = std::move(u);
                simplified = true;
            }
            return simplified;
        }
        case NodeKind::Internal: {
            if (0 == node.occupied_bits || node_children[node_key].kind == ChildrenKind::NoChildren) {
                nodes.item[node_key] = Node();
                return true;
            }
            if (node_children[node_key].kind != ChildrenKind::Children) return false;
            uint32_t child_keys[8];
            for (int i = 0; i < 8; ++i) child_keys[i] = node_children[node_key].child[i];

            simplify(child_keys[0]);
            if (!nodes.key_is_valid(child_keys[0])) {
                for (int i = 1; i < 8; ++i) simplify(child_keys[i]);
                return false;
            }
            for (int octant = 1; octant < 8; ++octant) {
                simplify(child_keys[octant]);
                if (!nodes.key_is_valid(child_keys[octant]) ||
                    !node_compare(nodes.item[child_keys[0]], nodes.item[child_keys[octant]]))
                    return false;
            }
            nodes.swap_items(node_key, child_keys[0]);
            const Children new_node_children = node_children[child_keys[0]];
            deallocate_children_of(node_key);
            node_children[node_key] = new_node_children;
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// octree/update/insert.rs:99-388
// ---------------------------------------------------------------------------------------------
static bool entry_is_none(const Entry& e) {  // mod.rs:102-109
    switch (e.kind) {
        case EntryKind::Empty: return true;
        case EntryKind::Visual: return e.albedo.a == 0;
        case EntryKind::Informative: return e.data == 0;
        case EntryKind::Complex: return e.albedo.a == 0 && e.data == 0;
    }
    return true;
}

Status Octree::insert_at_lod_internal(bool overwrite_if_empty, V3u position_u, uint32_t insert_size, const Entry& data) {
    const Cube root_bounds{unit(0.0f), (float)octree_size};
    const V3f position = to_f32(position_u);
    if (!bound_contains(root_bounds, position)) return E_INVALID_POSITION;
    if (entry_is_none(data)) return OK;

    struct StackItem {
        uint32_t key;
        Cube bounds;
    };
    std::vector<StackItem> node_stack;
    node_stack.push_back({0u, root_bounds});
    size_t actual_update_size = 0;
    const uint32_t target_content = add_to_palette(data);
    const Luts& l = luts();
    // `position.into()` : V3c<f32> -> V3c<u32> by round (vector.rs:326-336); exact for integral inputs
    const V3u position_back = position_u;

    for (;;) {
        const size_t current_node_key = node_stack.back().key;
        const Cube current_bounds = node_stack.back().bounds;
        const uint8_t target_child_octant = child_octant_for(current_bounds, position);
        const Cube target_bounds{
            current_bounds.min_position + l.octant_offset[target_child_octant] * current_bounds.size / 2.0f,
            current_bounds.size / 2.0f};

        size_t target_child_key = child_of(node_children[current_node_key], target_child_octant);
        if (insert_size > 1 && target_bounds.size <= (float)insert_size && lex_le(position, target_bounds.min_position)) {
            const NodeKind k = nodes.item[current_node_key].kind;
            if (k == NodeKind::Leaf || k == NodeKind::UniformLeaf) {
                subdivide_leaf_to_nodes(current_node_key, target_child_octant);
                target_child_key = child_of(node_children[current_node_key], target_child_octant);
            }
            if (nodes.key_is_valid(target_child_key)) {
                deallocate_children_of(target_child_key);
                Node n;
                n.kind = NodeKind::UniformLeaf;
                n.ubrick.kind = BrickKind::Solid;
                n.ubrick.solid = target_content;
                nodes.item[target_child_key] = std::move(n);
                Children& c = node_children[target_child_key];
                c = Children();
                c.kind = ChildrenKind::OccupancyBitmap;
                c.bitmap = UINT64_MAX;
            } else {
                Node n;
                n.kind = NodeKind::UniformLeaf;
                n.ubrick.kind = BrickKind::Solid;
                n.ubrick.solid = target_content;
                const uint32_t new_child_index = (uint32_t)nodes.push(std::move(n));
                if (node_children.size() < (size_t)new_child_index + 1) node_children.resize((size_t)new_child_index + 1);
                *child_mut(node_children[current_node_key], target_child_octant) = new_child_index;
                Children& c = node_children[new_child_index];
                c = Children();
                c.kind = ChildrenKind::OccupancyBitmap;
                c.bitmap = UINT64_MAX;
            }
            actual_update_size = f2usize(target_bounds.size);
            break;
        }

        const NodeKind kind = nodes.item[current_node_key].kind;
        const Cube data_bounds = (kind == NodeKind::UniformLeaf) ? current_bounds : target_bounds;

        if (data_bounds.size > (float)brick_dim || nodes.key_is_valid(target_child_key)) {
            if (nodes.key_is_valid(target_child_key)) {
                node_stack.push_back({(uint32_t)child_of(node_children[current_node_key], target_child_octant), target_bounds});
            } else {
                if (kind == NodeKind::Leaf || kind == NodeKind::UniformLeaf) {
                    const Node& cn = nodes.item[current_node_key];
                    bool target_match = false;
                    if (kind == NodeKind::UniformLeaf) {
                        const Brick& b = cn.ubrick;
                        if (b.kind == BrickKind::Solid)
                            target_match = (b.solid == target_content);
                        else if (b.kind == BrickKind::Parted) {
                            const V3s mi = matrix_index_for(current_bounds, position_back, brick_dim);
                            target_match = b.data[flat_projection(mi.x, mi.y, mi.z, brick_dim)] == target_content;
                        }
                    } else {
                        const Brick& b = cn.bricks[target_child_octant];
                        if (b.kind == BrickKind::Solid)
                            target_match = (b.solid == target_content);
                        else if (b.kind == BrickKind::Parted) {
                            const V3s mi = matrix_index_for(target_bounds, position_back, brick_dim);
                            target_match = b.data[flat_projection(mi.x, mi.y, mi.z, brick_dim)] == target_content;
                        }
                    }
                    if (target_match || node_is_all(cn, target_content)) break;

                    subdivide_leaf_to_nodes(current_node_key, target_child_octant);
                    node_stack.push_back(
                        {(uint32_t)child_of(node_children[current_node_key], target_child_octant), target_bounds});
                } else {
                    if (kind == NodeKind::Nothing) {
                        nodes.item[current_node_key].kind = NodeKind::Internal;
                        nodes.item[current_node_key].occupied_bits = 0;
                    }
                    const uint32_t new_child_node = (uint32_t)nodes.push(Node());
                    if (node_children.size() < nodes.len()) node_children.resize(nodes.len());
                    *child_mut(node_children[current_node_key], target_child_octant) = new_child_node;
                    node_stack.push_back({new_child_node, target_bounds});
                }
            }
        } else {
            actual_update_size = leaf_update(overwrite_if_empty, current_node_key, current_bounds, target_bounds,
                                             target_child_octant, position_back, insert_size, target_content);
            break;
        }
    }

    // post-processing operations (insert.rs:317-386)
    bool simplifyable = auto_simplify;
    for (size_t i = node_stack.size(); i-- > 0;) {
        const size_t node_key = node_stack[i].key;
        const Cube node_bounds = node_stack[i].bounds;
        if (!nodes.key_is_valid(node_key)) continue;
        if (nodes.item[node_key].kind == NodeKind::Nothing) {
            nodes.item[node_key].kind = NodeKind::Internal;
            nodes.item[node_key].occupied_bits = 0;
        }
        uint64_t new_occupied_bits = stored_occupied_bits(node_key);
        if (f2usize(node_bounds.size) == actual_update_size) {
            new_occupied_bits = UINT64_MAX;
        } else {
            const V3s rel = to_usize(position - node_bounds.min_position);
            set_occupancy_in_bitmap_64bits(rel.x, rel.y, rel.z, actual_update_size, f2usize(node_bounds.size), true,
                                           &new_occupied_bits);
        }
        store_occupied_bits(node_key, new_occupied_bits);
        // update MIP maps (insert.rs:371); `position` is the caller's u32 position
        update_mip(node_key, node_bounds, position_u);

        const NodeKind k = nodes.item[node_key].kind;
        if (k == NodeKind::Leaf || k == NodeKind::UniformLeaf) {
            simplifyable = simplify(node_key);
            continue;
        }
        if (simplifyable) simplifyable = simplify(node_key);
    }
    return OK;
}

// ---------------------------------------------------------------------------------------------
// octree/update/clear.rs and the emptiness helpers it needs (node.rs:107-241, :470-515, detail.rs:181-316)
// ---------------------------------------------------------------------------------------------
// BrickData::is_empty_throughout, node.rs:107-178
bool Octree::brick_is_empty_throughout(const Brick& b, uint8_t octant) const {
    const size_t d = brick_dim;
    switch (b.kind) {
        case BrickKind::Empty: return true;
        case BrickKind::Solid: return pix_points_to_empty(b.solid);
        case BrickKind::Parted: {
            if (1 == d) return pix_points_to_empty(b.data[0]);
            const V3s off = to_usize(luts().octant_offset[octant]);
            if (2 == d) return pix_points_to_empty(b.data[flat_projection(off.x, off.y, off.z, 2)]);
            const size_t extent = d / 2;
            for (size_t x = off.x * extent; x < off.x * extent + extent; ++x)
                for (size_t y = off.y * extent; y < off.y * extent + extent; ++y)
                    for (size_t z = off.z * extent; z < off.z * extent + extent; ++z)
                        if (!pix_points_to_empty(b.data[flat_projection(x, y, z, d)])) return false;
            return true;
        }
    }
    return true;
}

// BrickData::is_part_empty_throughout, node.rs:184-241
bool Octree::brick_is_part_empty_throughout(const Brick& b, uint8_t part_octant, uint8_t target_octant) const {
    const size_t d = brick_dim;
    switch (b.kind) {
        case BrickKind::Empty: return true;
        case BrickKind::Solid: return pix_points_to_empty(b.solid);
        case BrickKind::Parted: {
            if (1 == d) return pix_points_to_empty(b.data[0]);
            if (2 == d) {
                const V3s off = to_usize(luts().octant_offset[part_octant]);
                return pix_points_to_empty(b.data[flat_projection(off.x, off.y, off.z, 2)]);
            }
            const float outer_extent = (float)d / 2.0f, inner_extent = (float)d / 4.0f;
            const V3s off =
                to_usize(luts().octant_offset[part_octant] * outer_extent + luts().octant_offset[target_octant] * inner_extent);
            const size_t n = f2usize(inner_extent);
            for (size_t x = 0; x < n; ++x)
                for (size_t y = 0; y < n; ++y)
                    for (size_t z = 0; z < n; ++z)
                        if (!pix_points_to_empty(b.data[flat_projection(off.x + x, off.y + y, off.z + z, d)])) return false;
            return true;
        }
    }
    return true;
}

// NodeContent::is_empty, node.rs:470-515
bool Octree::node_is_empty(const Node& n) const {
    auto brick_empty = [&](const Brick& b) {
        if (b.kind == BrickKind::Empty) return true;
        if (b.kind == BrickKind::Solid) return pix_points_to_empty(b.solid);
        for (uint32_t v : b.data)
            if (!pix_points_to_empty(v)) return false;
        return true;
    };
    switch (n.kind) {
        case NodeKind::UniformLeaf: return brick_empty(n.ubrick);
        case NodeKind::Leaf:
            for (int o = 0; o < 8; ++o)
                if (!brick_empty(n.bricks[o])) return false;
            return true;
        case NodeKind::Internal: return false;
        case NodeKind::Nothing: return true;
    }
    return true;
}

// node_empty_at, detail.rs:255-316
bool Octree::node_empty_at(size_t node_key, uint8_t target_octant) const {
    const Node& n = nodes.item[node_key];
    auto brick_empty = [&](const Brick& b) {
        if (b.kind == BrickKind::Empty) return true;
        if (b.kind == BrickKind::Solid) return pix_points_to_empty(b.solid);
        const uint32_t* h = get_homogeneous_data(b);
        return h ? pix_points_to_empty(*h) : false;
    };
    switch (n.kind) {
        case NodeKind::Nothing: return true;
        case NodeKind::Leaf: return brick_empty(n.bricks[target_octant]);
        case NodeKind::UniformLeaf: return brick_empty(n.ubrick);
        case NodeKind::Internal:
            for (uint8_t child_octant = 0; child_octant < 8; ++child_octant) {
                const size_t child_key = child_of(node_children[node_key], target_octant);
                if (nodes.key_is_valid(child_key) && !node_empty_at(child_key, child_octant)) return false;
            }
            return true;
    }
    return true;
}

// should_bitmap_be_empty_at_bitmap_index -> ..._at_octants, detail.rs:181-252
bool Octree::should_bitmap_be_empty_at_bitmap_index(size_t node_key, size_t x, size_t y, size_t z) const {
    const V3f position = V3f{0.5f, 0.5f, 0.5f} + V3f{(float)x, (float)y, (float)z};
    const uint8_t target_octant = hash_region(position, (float)BITMAP_DIMENSION / 2.0f);
    const uint8_t target_octant_for_child =
        hash_region(position - (luts().octant_offset[target_octant] * (float)BITMAP_DIMENSION / 2.0f), (float)BITMAP_DIMENSION / 4.0f);
    const Node& n = nodes.item[node_key];
    switch (n.kind) {
        case NodeKind::Nothing: return true;
        case NodeKind::Internal: {
            const size_t child_key = child_of(node_children[node_key], target_octant);
            return nodes.key_is_valid(child_key) ? node_empty_at(child_key, target_octant_for_child) : true;
        }
        case NodeKind::UniformLeaf: return brick_is_part_empty_throughout(n.ubrick, target_octant, target_octant_for_child);
        case NodeKind::Leaf: return brick_is_empty_throughout(n.bricks[target_octant], target_octant_for_child);
    }
    return true;
}

// clear_at_lod, clear.rs:55-348
Status Octree::clear_at_lod(V3u position_u, uint32_t clear_size) {
    const Cube root_bounds{unit(0.0f), (float)octree_size};
    const V3f position = to_f32(position_u);
    if (!bound_contains(root_bounds, position)) return E_INVALID_POSITION;

    struct StackItem {
        uint32_t key;
        Cube bounds;
    };
    std::vector<StackItem> node_stack;
    node_stack.push_back({0u, root_bounds});
    size_t actual_update_size = 0;
    const Luts& l = luts();
    auto to_u32 = [](float v) { return (uint32_t)f2usize(std::round(v)); };  // From<V3c<f32>> for V3c<u32>, vector.rs:326-336

    for (;;) {
        const size_t current_node_key = node_stack.back().key;
        const Cube current_bounds = node_stack.back().bounds;
        const uint8_t target_child_octant = child_octant_for(current_bounds, position);
        const Cube target_bounds{
            current_bounds.min_position + l.octant_offset[target_child_octant] * current_bounds.size / 2.0f,
            current_bounds.size / 2.0f};
        const size_t target_child_key = child_of(node_children[current_node_key], target_child_octant);
        // `*position <= target_bounds.min_position.into()` : lexicographic compare of V3c<u32>
        const uint32_t mx = to_u32(target_bounds.min_position.x), my = to_u32(target_bounds.min_position.y),
                       mz = to_u32(target_bounds.min_position.z);
        const bool pos_le_min = position_u.x != mx ? position_u.x < mx : (position_u.y != my ? position_u.y < my : position_u.z <= mz);
        if (clear_size > 1 && target_bounds.size <= (float)clear_size && pos_le_min && nodes.key_is_valid(target_child_key)) {
            deallocate_children_of(target_child_key);
            nodes.item[target_child_key] = Node();
            node_children[target_child_key] = Children();
            actual_update_size = f2usize(target_bounds.size);
            node_stack.push_back({(uint32_t)child_of(node_children[current_node_key], target_child_octant), target_bounds});
            break;
        }

        if (target_bounds.size > (float)std::max(clear_size, brick_dim) || nodes.key_is_valid(target_child_key)) {
            if (nodes.key_is_valid(target_child_key)) {
                node_stack.push_back({(uint32_t)child_of(node_children[current_node_key], target_child_octant), target_bounds});
            } else {
                const Node& cn = nodes.item[current_node_key];
                if (cn.kind == NodeKind::Leaf || cn.kind == NodeKind::UniformLeaf) {
                    const Brick& b = cn.kind == NodeKind::UniformLeaf ? cn.ubrick : cn.bricks[target_child_octant];
                    bool target_match;
                    if (b.kind == BrickKind::Empty) {
                        target_match = true;
                    } else if (b.kind == BrickKind::Solid) {
                        target_match = pix_points_to_empty(b.solid);
                    } else {
                        // (sic) indexes the brick with position - current_bounds.min, unscaled (clear.rs:139-146, :164-171);
                        // the reference bounds-panics when that leaves the brick
                        const size_t ix = position_u.x - to_u32(current_bounds.min_position.x),
                                     iy = position_u.y - to_u32(current_bounds.min_position.y),
                                     iz = position_u.z - to_u32(current_bounds.min_position.z);
                        const size_t fi = flat_projection(ix, iy, iz, brick_dim);
                        if (fi >= b.data.size()) return E_INVALID_STRUCTURE;
                        target_match = pix_points_to_empty(b.data[fi]);
                    }
                    if (target_match || node_is_empty(cn)) break;
                    subdivide_leaf_to_nodes(current_node_key, target_child_octant);
                    node_stack.push_back({(uint32_t)child_of(node_children[current_node_key], target_child_octant), target_bounds});
                } else {
                    break;  // no child at the requested position: nothing to clear
                }
            }
        } else {
            actual_update_size = leaf_update(true, current_node_key, current_bounds, target_bounds, target_child_octant, position_u,
                                             clear_size, EMPTY_MARKER_U32);
            break;
        }
    }

    // post-processing (clear.rs:224-346)
    bool have_removed = false;
    StackItem removed{0, Cube{}};
    {
        const StackItem last = node_stack.back();
        node_stack.pop_back();
        if (f2usize(last.bounds.size) <= actual_update_size) {
            have_removed = true;
            removed = last;
        }
    }
    bool simplifyable = auto_simplify;
    for (size_t i = node_stack.size(); i-- > 0;) {
        const size_t node_key = node_stack[i].key;
        const Cube node_bounds = node_stack[i].bounds;
        if (have_removed) {
            const uint8_t child_octant = hash_region(
                (removed.bounds.min_position - node_bounds.min_position) + unit(removed.bounds.size / 2.0f), node_bounds.size / 2.0f);
            Children& c = node_children[node_key];
            if (c.kind == ChildrenKind::Children) {  // NodeChildren::clear, node.rs:75-83
                c.child[child_octant] = EMPTY_MARKER_U32;
                bool all_empty = true;
                for (int k = 0; k < 8; ++k) all_empty &= (c.child[k] == EMPTY_MARKER_U32);
                if (all_empty) c = Children();
            }
            nodes.free_key(removed.key);
            have_removed = false;
        }
        const uint64_t previous_occupied_bits = stored_occupied_bits(node_key);
        uint64_t new_occupied_bits = node_children[node_key].kind == ChildrenKind::NoChildren ? 0 : previous_occupied_bits;
        if (f2usize(node_bounds.size) == actual_update_size) {
            new_occupied_bits = 0;
        } else {
            const V3s start = matrix_index_for(node_bounds, position_u, (uint32_t)BITMAP_DIMENSION);
            const size_t n = f2usize(std::ceil((float)actual_update_size * (float)BITMAP_DIMENSION / node_bounds.size));
            for (size_t x = start.x; x < std::min(start.x + n, BITMAP_DIMENSION); ++x)
                for (size_t y = start.y; y < std::min(start.y + n, BITMAP_DIMENSION); ++y)
                    for (size_t z = start.z; z < std::min(start.z + n, BITMAP_DIMENSION); ++z)
                        if (should_bitmap_be_empty_at_bitmap_index(node_key, x, y, z))
                            set_occupancy_in_bitmap_64bits(x, y, z, 1, BITMAP_DIMENSION, false, &new_occupied_bits);
        }
        if (0 != new_occupied_bits && node_children[node_key].kind == ChildrenKind::Children) {
            Node in;
            in.kind = NodeKind::Internal;
            in.occupied_bits = new_occupied_bits;
            nodes.item[node_key] = std::move(in);
        } else {
            deallocate_children_of(node_key);
            node_children[node_key] = Children();
            have_removed = true;
            removed = node_stack[i];
            nodes.item[node_key] = Node();
        }
        if (0 == new_occupied_bits)
            node_children[node_key] = Children();
        else
            store_occupied_bits(node_key, new_occupied_bits);
        update_mip(node_key, node_bounds, position_u);  // clear.rs:335
        if (simplifyable) simplifyable = simplify(node_key);
        if (previous_occupied_bits == new_occupied_bits) break;
    }
    return OK;
}


// ---------------------------------------------------------------------------------------------
// octree/mipmap.rs
// ---------------------------------------------------------------------------------------------
// MIPMapStrategy::default(), mipmap.rs:591-604
MipStrategy::MipStrategy() {
    enabled = false;
    resampling_methods = {{1, {MipMethod::PointFilter, 0.0f}},
                          {2, {MipMethod::BoxFilter, 0.0f}},
                          {3, {MipMethod::BoxFilter, 0.0f}},
                          {4, {MipMethod::BoxFilter, 0.0f}}};
    resampling_color_matching_thresholds = {{2, 0.1f}, {3, 0.05f}, {4, 0.02f}};
}

// mipmap.rs:617-630
void Octree::mip_set_color_similarity_thr_at(size_t mip_level, float thr) {
    mip_map_strategy.resampling_color_matching_thresholds[mip_level] = clampf(thr, 0.0f, 1.0f);
}
// mipmap.rs:610-615
float Octree::mip_get_new_color_similarity_at(size_t mip_level) const {
    auto it = mip_map_strategy.resampling_color_matching_thresholds.find(mip_level);
    return it == mip_map_strategy.resampling_color_matching_thresholds.end() ? 0.0f : it->second;
}
// mipmap.rs:657-672
void Octree::mip_set_method_at(size_t mip_level, MipSampler method) {
    if (method.method == MipMethod::Posterize || method.method == MipMethod::PosterizeBD)
        method.thr = clampf(method.thr, 0.0f, 1.0f);
    else
        method.thr = 0.0f;
    mip_map_strategy.resampling_methods[mip_level] = method;
}
// mipmap.rs:650-655
MipSampler Octree::mip_get_method_at(size_t mip_level) const {
    auto it = mip_map_strategy.resampling_methods.find(mip_level);
    return it == mip_map_strategy.resampling_methods.end() ? MipSampler{} : it->second;
}

namespace {
// Albedou32, mipmap.rs:36-121. The reference computes in u32; in a release build `-` and `pow` wrap, and
// (2^32 - d)^2 mod 2^32 == d^2, so the wrapped distance below is the plain Euclidean one (a debug build panics).
struct Albedou32 {
    uint32_t r, g, b, a;
    bool operator==(const Albedou32& o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
};
inline Albedou32 au_from(Albedo c) { return {c.r, c.g, c.b, c.a}; }
inline Albedou32 au_pow2(Albedou32 c) { return {c.r * c.r, c.g * c.g, c.b * c.b, c.a * c.a}; }
inline Albedou32 au_add(Albedou32 x, Albedou32 y) { return {x.r + y.r, x.g + y.g, x.b + y.b, x.a + y.a}; }
inline Albedou32 au_sub(Albedou32 x, Albedou32 y) { return {x.r - y.r, x.g - y.g, x.b - y.b, x.a - y.a}; }
inline uint32_t f2u32(float v) {  // `f32 as u32`: saturating, NaN -> 0
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
inline Albedou32 au_div(Albedou32 c, uint32_t d) {  // :88-98
    return {f2u32(std::round((float)c.r / (float)d)), f2u32(std::round((float)c.g / (float)d)),
            f2u32(std::round((float)c.b / (float)d)), f2u32(std::round((float)c.a / (float)d))};
}
inline Albedou32 au_sqrt(Albedou32 c) {  // :47-53
    return {f2u32(std::round(std::sqrt((float)c.r))), f2u32(std::round(std::sqrt((float)c.g))),
            f2u32(std::round(std::sqrt((float)c.b))), f2u32(std::round(std::sqrt((float)c.a)))};
}
inline float au_length(Albedou32 c) {  // :44-46
    return std::sqrt((float)(uint32_t)(c.r * c.r + c.g * c.g + c.b * c.b + c.a * c.a));
}
inline Albedo au_to_albedo(Albedou32 c) {  // :112-121
    return {(uint8_t)std::min(c.r, 255u), (uint8_t)std::min(c.g, 255u), (uint8_t)std::min(c.b, 255u),
            (uint8_t)std::min(c.a, 255u)};
}
inline uint8_t f2u8(float v) {  // `f32 as u8`
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
inline bool albedo_eq(Albedo x, Albedo y) { return x.r == y.r && x.g == y.g && x.b == y.b && x.a == y.a; }

// MIPResaplingFunction::execute, mipmap.rs:133-263. `sample_fn` returns false for None.
//
// DETERMINISM NOTE. PointFilter and Posterize keep their candidates in a std HashMap and the reference takes
// `into_iter().max_by_key(count)` (the LAST maximum in iteration order) - with Rust's randomly seeded hasher a tie
// between two colours is resolved differently from run to run, and Posterize's `for .. in albedo_counts.iter()` picks
// "the first bucket within the threshold" in that same random order. Any of those outcomes is a valid reference
// result; this restatement (and the product) fix the order to FIRST-SEEN: buckets are kept in the order their first
// sample arrived (x outer, y, z inner), a re-keyed Posterize bucket keeps its place, and a tie goes to the earliest.
template <typename F>
bool mip_sample(const MipSampler& sampler, V3u sample_start, uint32_t sample_size, F&& sample_fn, Albedo* out) {
    switch (sampler.method) {
        case MipMethod::BoxFilter: {
            bool have = false;
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int32_t entry_count = 0;
            for (uint32_t x = sample_start.x; x < sample_start.x + sample_size; ++x)
                for (uint32_t y = sample_start.y; y < sample_start.y + sample_size; ++y)
                    for (uint32_t z = sample_start.z; z < sample_start.z + sample_size; ++z) {
                        Albedo c;
                        if (!sample_fn(V3u{x, y, z}, &c)) continue;
                        if (!have) {
                            have = true;
                            entry_count = 1;
                            s0 = (float)c.r * (float)c.r;
                            s1 = (float)c.g * (float)c.g;
                            s2 = (float)c.b * (float)c.b;
                            s3 = (float)c.a * (float)c.a;
                        } else {
                            entry_count += 1;
                            s0 += (float)c.r * (float)c.r;
                            s1 += (float)c.g * (float)c.g;
                            s2 += (float)c.b * (float)c.b;
                            s3 += (float)c.a * (float)c.a;
                        }
                    }
            if (!have) return false;
            out->r = f2u8(fmin_(std::sqrt(s0 / (float)entry_count), 255.0f));
            out->g = f2u8(fmin_(std::sqrt(s1 / (float)entry_count), 255.0f));
            out->b = f2u8(fmin_(std::sqrt(s2 / (float)entry_count), 255.0f));
            out->a = f2u8(fmin_(std::sqrt(s3 / (float)entry_count), 255.0f));
            return true;
        }
        case MipMethod::PointFilter:
        case MipMethod::PointFilterBD: {
            std::vector<std::pair<Albedo, uint32_t>> counts;
            for (uint32_t x = sample_start.x; x < sample_start.x + sample_size; ++x)
                for (uint32_t y = sample_start.y; y < sample_start.y + sample_size; ++y)
                    for (uint32_t z = sample_start.z; z < sample_start.z + sample_size; ++z) {
                        Albedo c;
                        if (!sample_fn(V3u{x, y, z}, &c)) continue;
                        bool found = false;
                        for (auto& e : counts)
                            if (albedo_eq(e.first, c)) {
                                e.second += 1;
                                found = true;
                                break;
                            }
                        if (!found) counts.push_back({c, 1});
                    }
            if (counts.empty()) return false;
            size_t best = 0;
            for (size_t i = 1; i < counts.size(); ++i)
                if (counts[i].second > counts[best].second) best = i;
            *out = counts[best].first;
            return true;
        }
        case MipMethod::Posterize:
        case MipMethod::PosterizeBD: {
            const float thr = sampler.thr;
            std::vector<std::pair<Albedou32, uint32_t>> counts;  // squared sums, occurrence counts
            for (uint32_t x = sample_start.x; x < sample_start.x + sample_size; ++x)
                for (uint32_t y = sample_start.y; y < sample_start.y + sample_size; ++y)
                    for (uint32_t z = sample_start.z; z < sample_start.z + sample_size; ++z) {
                        Albedo c;
                        if (!sample_fn(V3u{x, y, z}, &c)) continue;
                        size_t hit = counts.size();
                        for (size_t i = 0; i < counts.size(); ++i) {
                            const Albedou32 poster_color = au_sqrt(au_div(counts[i].first, counts[i].second));
                            if (au_length(au_sub(poster_color, au_from(c))) < (thr * 255.0f)) {
                                hit = i;
                                break;
                            }
                        }
                        if (hit < counts.size()) {
                            // remove(old key) + insert(new key, count + 1); an insert onto an existing key replaces
                            // that entry's count (HashMap::insert) - the colliding bucket is dropped
                            const Albedou32 new_sum = au_add(counts[hit].first, au_pow2(au_from(c)));
                            const uint32_t new_count = counts[hit].second + 1;
                            counts[hit] = {new_sum, new_count};
                            for (size_t i = 0; i < counts.size(); ++i)
                                if (i != hit && counts[i].first == new_sum) {
                                    counts.erase(counts.begin() + (ptrdiff_t)i);
                                    break;
                                }
                        } else {
                            const Albedou32 key = au_pow2(au_from(c));
                            bool replaced = false;
                            for (auto& e : counts)
                                if (e.first == key) {
                                    e.second = 1;
                                    replaced = true;
                                    break;
                                }
                            if (!replaced) counts.push_back({key, 1});
                        }
                    }
            if (counts.empty()) return false;
            size_t best = 0;
            for (size_t i = 1; i < counts.size(); ++i)
                if (counts[i].second > counts[best].second) best = i;
            *out = au_to_albedo(au_sqrt(au_div(counts[best].first, counts[best].second)));
            return true;
        }
    }
    return false;
}

inline bool entry_albedo(const Entry& e, Albedo* out) {  // OctreeEntry::albedo, mod.rs:79-86
    if (e.kind == EntryKind::Visual || e.kind == EntryKind::Complex) {
        *out = e.albedo;
        return true;
    }
    return false;
}
// Albedo::distance_from, detail.rs:82-89
inline float albedo_distance(Albedo x, Albedo y) {
    const float dr = (float)x.r - (float)y.r, dg = (float)x.g - (float)y.g, db = (float)x.b - (float)y.b,
                da = (float)x.a - (float)y.a;
    return std::sqrt(dr * dr + dg * dg + db * db + da * da);
}
inline uint32_t f2u32_round(float v) { return f2u32(std::round(v)); }  // From<V3c<f32>> for V3c<u32>, vector.rs:326-336
}  // namespace

// mipmap.rs:296-584
void Octree::update_mip(size_t node_key, const Cube& node_bounds, V3u position) {
    if (!mip_map_strategy.enabled) return;
    ensure_mips();
    const size_t mip_level = f2usize(std::log2(node_bounds.size / (float)brick_dim));
    MipSampler sampler;  // MIPResamplingMethods::default() == BoxFilter
    {
        auto it = mip_map_strategy.resampling_methods.find(mip_level);
        if (it != mip_map_strategy.resampling_methods.end()) sampler = it->second;
    }
    const bool dominant_bottom = sampler.method == MipMethod::PointFilterBD;  // :309-314: PosterizeBD is NOT matched here

    const NodeKind kind = nodes.item[node_key].kind;
    V3u sample_start{0, 0, 0};
    uint32_t sample_size = 0;
    const uint32_t size_u = f2u32(node_bounds.size);
    switch (kind) {
        case NodeKind::Nothing: return;
        case NodeKind::UniformLeaf:
            // Uniform leaf nodes need no MIP, their content is equivalent with it (:331-337)
            node_mips[node_key] = Brick();
            return;
        case NodeKind::Leaf: {
            sample_size = std::min(size_u / brick_dim, brick_dim * 2);
            // insert_at_lod with a size <= brick_dim can split leaves into nodes SMALLER than a brick (insert.rs:137-147);
            // for those `size / dim` is 0 and the reference panics on `position % 0` - the MIP is left alone here
            if (sample_size == 0) return;
            const V3u t = {(position.x - (position.x % sample_size)) * 2 * brick_dim,
                           (position.y - (position.y % sample_size)) * 2 * brick_dim,
                           (position.z - (position.z % sample_size)) * 2 * brick_dim};
            const V3f f = to_f32(t) / node_bounds.size;
            sample_start = {f2u32_round(std::floor(f.x)), f2u32_round(std::floor(f.y)), f2u32_round(std::floor(f.z))};
            break;
        }
        case NodeKind::Internal:
            if (dominant_bottom) {
                sample_size = size_u / brick_dim;
                if (sample_size == 0) return;  // as above: remainder by zero, a panic in the reference
                const V3f f = to_f32(V3u{position.x - (position.x % sample_size), position.y - (position.y % sample_size),
                                         position.z - (position.z % sample_size)});
                sample_start = {f2u32_round(std::floor(f.x)), f2u32_round(std::floor(f.y)), f2u32_round(std::floor(f.z))};
            } else {
                sample_size = 2;
                const V3f pos_in_bounds = to_f32(position) - node_bounds.min_position;
                const V3f f = pos_in_bounds * 2.0f * (float)brick_dim / node_bounds.size;  // into 2*DIM space
                sample_start = {f2u32_round(std::floor(f.x)), f2u32_round(std::floor(f.y)), f2u32_round(std::floor(f.z))};
                sample_start = {sample_start.x - (sample_start.x % 2), sample_start.y - (sample_start.y % 2),
                                sample_start.z - (sample_start.z % 2)};
            }
            break;
    }

    auto sample_tree = [&](V3u pos, Albedo* out) { return entry_albedo(get_internal(node_key, node_bounds, pos), out); };
    Albedo sampled_color{0, 0, 0, 0};
    bool have_color = false;
    if (kind == NodeKind::Leaf || (kind == NodeKind::Internal && dominant_bottom)) {
        have_color = mip_sample(sampler, sample_start, sample_size, sample_tree, &sampled_color);
    } else {  // Internal: sample the MIPs of the children; the range spans 0 .. 2*brick_dim (:458-511)
        const Children& ch = node_children[node_key];
        if ((size_t)EMPTY_MARKER_U32 != child_of(ch, hash_region(to_f32(sample_start), (float)brick_dim))) {
            auto sample_child_mips = [&](V3u pos, Albedo* out) -> bool {
                const uint8_t child_octant = hash_region(to_f32(pos), (float)brick_dim);
                const size_t child_key = child_of(ch, child_octant);
                if ((size_t)EMPTY_MARKER_U32 == child_key) return false;
                if (child_key >= node_mips.size()) return false;  // the reference would index out of bounds and panic
                const V3f off = luts().octant_offset[child_octant] * (float)brick_dim;
                const V3u p = {pos.x - f2u32_round(off.x), pos.y - f2u32_round(off.y), pos.z - f2u32_round(off.z)};
                const Brick& m = node_mips[child_key];
                switch (m.kind) {
                    case BrickKind::Empty: return false;
                    case BrickKind::Solid: return entry_albedo(pix_get_ref(m.solid), out);
                    case BrickKind::Parted: return entry_albedo(pix_get_ref(m.data[flat_projection(p.x, p.y, p.z, brick_dim)]), out);
                }
                return false;
            };
            have_color = mip_sample(sampler, sample_start, sample_size, sample_child_mips, &sampled_color);
        }
    }
    if (!have_color) return;  // a MIP entry is never cleared (:513-548 only ever writes Some)

    // Assemble MIP entry (:513-548)
    Entry visual;
    visual.kind = EntryKind::Visual;
    visual.albedo = sampled_color;
    uint32_t mip_entry;
    auto thr_it = mip_map_strategy.resampling_color_matching_thresholds.find(mip_level);
    if (thr_it != mip_map_strategy.resampling_color_matching_thresholds.end()) {
        const float color_distance_threshold = thr_it->second * 255.0f;
        bool found = false;
        size_t similar = 0;
        for (size_t i = 0; i < voxel_color_palette.size(); ++i)
            if (albedo_distance(sampled_color, voxel_color_palette[i]) < color_distance_threshold) {
                similar = i;
                found = true;
                break;
            }
        mip_entry = found ? pix_visual((uint16_t)similar) : add_to_palette(visual);
    } else {
        mip_entry = add_to_palette(visual);
    }

    // Set MIP entry (:550-581)
    const V3s pim = matrix_index_for(node_bounds, position, brick_dim);
    const size_t flat = flat_projection(pim.x, pim.y, pim.z, brick_dim);
    Brick& mip = node_mips[node_key];
    const size_t volume = (size_t)brick_dim * brick_dim * brick_dim;
    switch (mip.kind) {
        case BrickKind::Empty:
            mip.data.assign(volume, EMPTY_MARKER_U32);
            mip.kind = BrickKind::Parted;
            break;
        case BrickKind::Solid:
            mip.data.assign(volume, mip.solid);
            mip.kind = BrickKind::Parted;
            break;
        case BrickKind::Parted: break;
    }
    mip.witness = 0;
    mip.data.forget();  // written in place below
    if (flat < mip.data.size()) mip.data[flat] = mip_entry;
}

// mipmap.rs:875-892
void Octree::recalculate_mip(size_t node_key, const Cube& node_bounds) {
    if (!mip_map_strategy.enabled) return;
    for (uint32_t x = 0; x < brick_dim; ++x)
        for (uint32_t y = 0; y < brick_dim; ++y)
            for (uint32_t z = 0; z < brick_dim; ++z) {
                const V3f o = V3f{(float)x, (float)y, (float)z} * node_bounds.size / (float)brick_dim;
                const V3f pos = node_bounds.min_position + V3f{std::round(o.x), std::round(o.y), std::round(o.z)};
                update_mip(node_key, node_bounds, V3u{f2u32_round(pos.x), f2u32_round(pos.y), f2u32_round(pos.z)});
            }
}

// mipmap.rs:798-855: depth first, children (ascending octant) before their parent
void Octree::recalculate_mips() {
    node_mips.assign(nodes.len(), Brick());
    struct Frame {
        size_t key;
        Cube bounds;
        uint8_t target_octant;
    };
    std::vector<Frame> node_stack;
    node_stack.push_back({0, Cube{unit(0.0f), (float)octree_size}, 0});
    while (!node_stack.empty()) {
        Frame& top = node_stack.back();
        if (OOB_OCTANT == top.target_octant) {
            const size_t key = top.key;
            const Cube bounds = top.bounds;
            recalculate_mip(key, bounds);
            node_stack.pop_back();
            if (!node_stack.empty()) node_stack.back().target_octant += 1;
            continue;
        }
        switch (nodes.item[top.key].kind) {
            case NodeKind::Nothing:
                // unreachable!() in the reference: only a Nothing ROOT can get here and the public entry point
                // (switch_albedo_mip_maps) checks for it; recalculate_mips() on an empty tree would panic
                node_stack.pop_back();
                break;
            case NodeKind::Internal: {
                const size_t child = child_of(node_children[top.key], top.target_octant);
                if (nodes.key_is_valid(child) && nodes.item[child].kind != NodeKind::Nothing) {
                    const Cube cb = child_bounds_for(top.bounds, top.target_octant);
                    node_stack.push_back({child, cb, 0});
                } else {
                    top.target_octant += 1;
                }
                break;
            }
            case NodeKind::Leaf:
            case NodeKind::UniformLeaf: top.target_octant = OOB_OCTANT; break;
        }
    }
}

// mipmap.rs:858-872
void Octree::switch_albedo_mip_maps(bool enabled) {
    const bool mips_on_previously = mip_map_strategy.enabled;
    mip_map_strategy.enabled = enabled;
    if (mip_map_strategy.enabled && mips_on_previously != enabled && nodes.item[0].kind != NodeKind::Nothing)
        recalculate_mips();
}

// mipmap.rs:897-937
Entry Octree::sample_root_mip(uint8_t octant, V3u position) const {
    const size_t node_key = OOB_OCTANT == octant ? 0 : child_of(node_children[0], octant);
    if (!nodes.key_is_valid(node_key) || node_key >= node_mips.size()) return Entry();
    const Brick& m = node_mips[node_key];
    switch (m.kind) {
        case BrickKind::Empty: return Entry();
        case BrickKind::Solid: return pix_get_ref(m.solid);
        case BrickKind::Parted: return pix_get_ref(m.data[flat_projection(position.x, position.y, position.z, brick_dim)]);
    }
    return Entry();
}

// ---------------------------------------------------------------------------------------------
// structure hash: key-order independent digest of everything get_by_ray can observe
// ---------------------------------------------------------------------------------------------
static inline uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    return h;
}
static uint64_t hash_brick(const Brick& b) {
    uint64_t h = mix64(0x1234, (uint64_t)b.kind);
    if (b.kind == BrickKind::Solid) h = mix64(h, b.solid);
    if (b.kind == BrickKind::Parted)
        for (uint32_t v : b.data) h = mix64(h, v);
    return h;
}
uint64_t Octree::hash_node(size_t key) const {
    const Node& n = nodes.item[key];
    uint64_t h = mix64(0xABCD, (uint64_t)n.kind);
    h = mix64(h, stored_occupied_bits(key));
    switch (n.kind) {
        case NodeKind::Nothing: break;
        case NodeKind::Internal:
            for (uint8_t o = 0; o < 8; ++o) {
                const size_t c = child_of(node_children[key], o);
                h = mix64(h, nodes.key_is_valid(c) ? hash_node(c) : 0x5EED);
            }
            break;
        case NodeKind::Leaf:
            for (int o = 0; o < 8; ++o) h = mix64(h, hash_brick(n.bricks[o]));
            break;
        case NodeKind::UniformLeaf: h = mix64(h, hash_brick(n.ubrick)); break;
    }
    return h;
}
uint64_t Octree::structure_hash() const {
    uint64_t h = mix64(octree_size, brick_dim);
    for (const Albedo& a : voxel_color_palette)
        h = mix64(h, ((uint64_t)a.r << 24) | ((uint64_t)a.g << 16) | ((uint64_t)a.b << 8) | a.a);
    h = mix64(h, 0xDA7A);
    for (uint32_t d : voxel_data_palette) h = mix64(h, d);
    return mix64(h, hash_node(0));
}

// MIP digest: the strategy plus the MIP brick of every reachable node, in the same traversal order as hash_node
static uint64_t mip_hash_node(const Octree& t, size_t key) {
    static const Brick none;
    uint64_t h = mix64(0x313D, hash_brick(key < t.node_mips.size() ? t.node_mips[key] : none));
    if (t.nodes.item[key].kind == NodeKind::Internal)
        for (uint8_t o = 0; o < 8; ++o) {
            const size_t c = child_of(t.node_children[key], o);
            h = mix64(h, t.nodes.key_is_valid(c) ? mip_hash_node(t, c) : 0x5EED);
        }
    return h;
}
uint64_t Octree::mip_hash() const {
    uint64_t h = mix64(0x57A7, mip_map_strategy.enabled ? 1 : 0);
    for (const auto& m : mip_map_strategy.resampling_methods) {
        uint32_t bits;
        std::memcpy(&bits, &m.second.thr, 4);
        h = mix64(mix64(mix64(h, m.first), (uint64_t)m.second.method), bits);
    }
    h = mix64(h, 0x7447);
    for (const auto& m : mip_map_strategy.resampling_color_matching_thresholds) {
        uint32_t bits;
        std::memcpy(&bits, &m.second, 4);
        h = mix64(mix64(h, m.first), bits);
    }
    return mix64(h, mip_hash_node(*this, 0));
}

// ---------------------------------------------------------------------------------------------
// raytracing/raytracing_on_cpu.rs
// ---------------------------------------------------------------------------------------------
// :99-112 ; `x.powf(2.)` is evaluated as x*x (LLVM folds pow(x, 2.0) -> x*x)
V3f get_dda_scale_factors(const Ray& ray) {
    const V3f d = ray.direction;
    auto sq = [](float v) { return v * v; };
    return {std::sqrt(1.0f + sq(d.z / d.x) + sq(d.y / d.x)), std::sqrt(sq(d.x / d.y) + 1.0f + sq(d.z / d.y)),
            std::sqrt((sq(d.x / d.z) + 1.0f) + sq(d.y / d.z))};
}

// :124-152
V3f dda_step_to_next_sibling(const Ray& ray, V3f& p, const Cube& current_bounds, const V3f& scale) {
    const V3f sg = {signum(ray.direction.x), signum(ray.direction.y), signum(ray.direction.z)};
    const V3f diff_from_min = p - current_bounds.min_position;
    const V3f steps_needed = {current_bounds.size * fmax_(sg.x, 0.0f) - sg.x * diff_from_min.x,
                              current_bounds.size * fmax_(sg.y, 0.0f) - sg.y * diff_from_min.y,
                              current_bounds.size * fmax_(sg.z, 0.0f) - sg.z * diff_from_min.z};
    const float d_x = std::fabs(steps_needed.x * scale.x);
    const float d_y = std::fabs(steps_needed.y * scale.y);
    const float d_z = std::fabs(steps_needed.z * scale.z);
    const float min_step = fmin_(fmin_(d_x, d_y), d_z);
    p = p + ray.direction * min_step;
    return {min_step == d_x ? sg.x : 0.0f, min_step == d_y ? sg.y : 0.0f, min_step == d_z ? sg.z : 0.0f};
}

// :156-252
bool Octree::traverse_brick(const Ray& ray, V3f& p, const std::vector<uint32_t>& brick, const Cube& brick_bounds,
                            const V3f& scale, int32_t idx_out[3], size_t& flat_out, RayStats* st) const {
    const int32_t dim = (int32_t)brick_dim;
    const V3f pos_in_brick = (p - brick_bounds.min_position) * (float)brick_dim / brick_bounds.size;
    int32_t ix = std::clamp(f2i32(pos_in_brick.x), 0, dim - 1);
    int32_t iy = std::clamp(f2i32(pos_in_brick.y), 0, dim - 1);
    int32_t iz = std::clamp(f2i32(pos_in_brick.z), 0, dim - 1);
    const int32_t fdx = 1, fdy = dim, fdz = dim * dim;
    int32_t flat = (int32_t)flat_projection((size_t)ix, (size_t)iy, (size_t)iz, brick_dim);

    const float brick_unit = brick_bounds.size / (float)brick_dim;
    Cube current_bounds{brick_bounds.min_position + V3f{(float)ix, (float)iy, (float)iz} * brick_unit, brick_unit};

    V3f step = unit(0.0f);
    for (;;) {
        if (ix < 0 || ix >= dim || iy < 0 || iy >= dim || iz < 0 || iz >= dim) return false;
        flat += f2i32(step.x) * fdx + f2i32(step.y) * fdy + f2i32(step.z) * fdz;
        if (st) st->voxel_fetches++;
        if (!pix_points_to_empty(brick[(size_t)flat])) {
            idx_out[0] = ix;
            idx_out[1] = iy;
            idx_out[2] = iz;
            flat_out = (size_t)flat;
            return true;
        }
        step = dda_step_to_next_sibling(ray, p, current_bounds, scale);
        current_bounds.min_position = current_bounds.min_position + step * brick_unit;
        // V3c::<i32>::from(step): round() as i32 (vector.rs:354-364)
        ix += f2i32(std::round(step.x));
        iy += f2i32(std::round(step.y));
        iz += f2i32(std::round(step.z));
    }
}

// :256-312
bool Octree::probe_brick(const Ray& ray, V3f& p, const Brick& brick, const Cube& brick_bounds, const V3f& scale, Hit& out,
                         RayStats* st) const {
    switch (brick.kind) {
        case BrickKind::Empty: return false;
        case BrickKind::Solid:
            out.hit = true;
            out.palette_value = brick.solid;
            out.entry = pix_get_ref(brick.solid);
            out.impact_point = p;
            out.normal = cube_impact_normal(brick_bounds, p);
            return true;
        case BrickKind::Parted: {
            int32_t idx[3];
            size_t flat;
            if (traverse_brick(ray, p, brick.data, brick_bounds, scale, idx, flat, st)) {
                // V3c::<usize>::from(current_index) then V3c::<f32>::from(..)
                const V3f idxf = {(float)(size_t)idx[0], (float)(size_t)idx[1], (float)(size_t)idx[2]};
                const Cube hit_bounds{brick_bounds.min_position + idxf * brick_bounds.size / (float)brick_dim,
                                      brick_bounds.size / (float)brick_dim};
                out.hit = true;
                out.palette_value = brick.data[flat];
                out.entry = pix_get_ref(brick.data[flat]);
                out.impact_point = p;
                out.normal = cube_impact_normal(hit_bounds, p);
                return true;
            }
            return false;
        }
    }
    return false;
}

// :316-318
Hit Octree::get_by_ray(const Ray& ray, RayStats* st) const { return get_by_ray_at_lod(ray, std::numeric_limits<float>::max(), st); }

// :325-565
Hit Octree::get_by_ray_at_lod(const Ray& ray, float viewing_distance, RayStats* st) const {
    Hit result;
    const Luts& l = luts();
    const V3f ray_scale_factors = get_dda_scale_factors(ray);
    const size_t direction_lut_index = hash_direction(ray.direction);

    NodeStack<uint32_t, 4> node_stack;
    Cube current_bounds{unit(0.0f), (float)octree_size};
    V3f ray_current_point;
    uint8_t target_octant;
    {
        bool has_d;
        float d;
        if (intersect_ray(current_bounds, ray, &has_d, &d)) {
            // ray.point_at(d) = origin + direction * d  (spatial/raytracing/mod.rs:18-20)
            ray_current_point = ray.origin + ray.direction * (has_d ? d : 0.0f);
            target_octant = hash_region(ray_current_point - current_bounds.min_position, current_bounds.size / 2.0f);
        } else {
            ray_current_point = ray.origin;
            target_octant = OOB_OCTANT;
        }
    }
    size_t current_node_key;
    // :349; never reset at a restart from the root and not decremented by the root push, so it drifts by +1 per
    // completed root cycle (and the other way when the 4-entry ring stack loses entries) - kept as is
    float mip_level = std::log2((float)octree_size / (float)brick_dim);

    auto bitmap_index = [&](const V3f& bp) -> size_t {
        // `f.floor() as usize` saturates; an index > 3 bounds-panics in the reference
        size_t bx = f2usize(std::floor(bp.x)), by = f2usize(std::floor(bp.y)), bz = f2usize(std::floor(bp.z));
        if (bx > 3 || by > 3 || bz > 3) {
            if (st) st->would_panic++;
            bx = std::min<size_t>(bx, 3);
            by = std::min<size_t>(by, 3);
            bz = std::min<size_t>(bz, 3);
        }
        return l.bitmap_index[bx][by][bz];
    };

    while (target_octant != OOB_OCTANT) {
        if (st) st->outer_iters++;
        const uint32_t node_iters_before = st ? st->node_iters : 0;
        current_node_key = 0;
        current_bounds = Cube{unit(0.0f), (float)octree_size};
        node_stack.push(0);
        while (!node_stack.is_empty()) {
            if (st) st->node_iters++;
            const uint64_t current_node_occupied_bits = stored_occupied_bits(*node_stack.last());
            const Node& cur = nodes.item[current_node_key];
            bool do_backtrack_after_leaf_miss = (cur.kind == NodeKind::UniformLeaf);

            // :368-386 the node's MIP stands in for its content once the ray has travelled far enough; a miss
            // leaves ray_current_point advanced by the brick walk and target_octant untouched
            if (mip_map_strategy.enabled) {
                const float m2 = mip_level * 2.0f;
                const V3f q = ray_current_point / m2;
                const V3f aligned = V3f{std::round(q.x), std::round(q.y), std::round(q.z)} * m2;
                if (mip_level < length(ray.origin - aligned) / viewing_distance) {
                    if (st) st->mip_probes++;
                    static const Brick no_mip;
                    const Brick& mip = current_node_key < node_mips.size() ? node_mips[current_node_key] : no_mip;
                    if (probe_brick(ray, ray_current_point, mip, current_bounds, ray_scale_factors, result, st))
                        return result;
                }
            }

            if (target_octant != OOB_OCTANT) {
                if (cur.kind == NodeKind::UniformLeaf) {
                    if (probe_brick(ray, ray_current_point, cur.ubrick, current_bounds, ray_scale_factors, result, st))
                        return result;
                    do_backtrack_after_leaf_miss = true;
                } else if (cur.kind == NodeKind::Leaf) {
                    if (probe_brick(ray, ray_current_point, cur.bricks[target_octant],
                                    child_bounds_for(current_bounds, target_octant), ray_scale_factors, result, st))
                        return result;
                }
            }

            V3f bitmap_pos_in_node =
                (ray_current_point - current_bounds.min_position) * (float)BITMAP_DIMENSION / current_bounds.size;
            bitmap_pos_in_node = {clampf(bitmap_pos_in_node.x, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE),
                                  clampf(bitmap_pos_in_node.y, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE),
                                  clampf(bitmap_pos_in_node.z, FLOAT_ERROR_TOLERANCE, 4.0f - FLOAT_ERROR_TOLERANCE)};
            size_t flat_pos_in_bitmap = bitmap_index(bitmap_pos_in_node);

            if (do_backtrack_after_leaf_miss || target_octant == OOB_OCTANT || 0 == current_node_occupied_bits ||
                0 == (current_node_occupied_bits & l.ray_to_node_occupancy[flat_pos_in_bitmap][direction_lut_index])) {
                // POP
                node_stack.pop(nullptr);
                mip_level += 1.0f;
                if (const uint32_t* parent = node_stack.last()) {
                    current_node_key = *parent;
                    const V3f current_bound_center = current_bounds.min_position + unit(current_bounds.size / 2.0f);
                    // V3c::modulo: per-component `%` (fmod) (vector.rs:66-71)
                    const float m = current_bounds.size * 2.0f;
                    const V3f mod = {std::fmod(current_bounds.min_position.x, m), std::fmod(current_bounds.min_position.y, m),
                                     std::fmod(current_bounds.min_position.z, m)};
                    const V3f parent_bound_min_position = current_bounds.min_position - mod;
                    const uint8_t from = hash_region(current_bound_center - parent_bound_min_position, current_bounds.size);
                    const V3f step = dda_step_to_next_sibling(ray, ray_current_point, current_bounds, ray_scale_factors);
                    target_octant = step_octant(from, step);
                    current_bounds.size *= 2.0f;
                    current_bounds.min_position = parent_bound_min_position;
                }
                continue;
            }

            Cube target_bounds = child_bounds_for(current_bounds, target_octant);
            uint32_t target_child_key = (uint32_t)child_of(node_children[current_node_key], target_octant);
            if (nodes.key_is_valid(target_child_key) &&
                0 != (current_node_occupied_bits & l.bitmap_mask_for_octant[target_octant])) {
                // PUSH
                current_node_key = target_child_key;
                current_bounds = target_bounds;
                target_octant = hash_region(ray_current_point - target_bounds.min_position, target_bounds.size / 2.0f);
                node_stack.push(target_child_key);
                mip_level -= 1.0f;
            } else {
                // ADVANCE
                for (;;) {
                    const V3f step_vec = dda_step_to_next_sibling(ray, ray_current_point, target_bounds, ray_scale_factors);
                    target_octant = step_octant(target_octant, step_vec);
                    if (OOB_OCTANT != target_octant) {
                        target_bounds = child_bounds_for(current_bounds, target_octant);
                        target_child_key = (uint32_t)child_of(node_children[current_node_key], target_octant);
                        bitmap_pos_in_node = bitmap_pos_in_node + step_vec * 4.0f / current_bounds.size;
                        flat_pos_in_bitmap = bitmap_index(bitmap_pos_in_node);
                    }
                    bool stop = (target_octant == OOB_OCTANT);
                    if (!stop)
                        stop = nodes.key_is_valid(target_child_key) &&
                               0 != (current_node_occupied_bits & l.bitmap_mask_for_octant[target_octant]) &&
                               0 != (l.ray_to_node_occupancy[flat_pos_in_bitmap][direction_lut_index] &
                                     current_node_occupied_bits);
                    if (!stop) {
                        const Node& cn = nodes.item[current_node_key];
                        if (cn.kind == NodeKind::Leaf) stop = cn.bricks[target_octant].kind != BrickKind::Empty;
                    }
                    if (stop) break;
                }
            }
        }
        if (st && st->node_iters - node_iters_before == 1) st->crawl_iters++;
        // :548-562 restart from the root after nudging the point forward
        ray_current_point = ray_current_point + ray.direction * 0.1f;
        const float sz = (float)octree_size;
        if (ray_current_point.x < sz && ray_current_point.y < sz && ray_current_point.z < sz &&
            ray_current_point.x > 0.0f && ray_current_point.y > 0.0f && ray_current_point.z > 0.0f)
            target_octant = hash_region(ray_current_point, sz / 2.0f);
        else
            target_octant = OOB_OCTANT;
    }
    return result;
}

// ---------------------------------------------------------------------------------------------
// caller-side ray generation, examples/cpu_render.rs:78-114 (same in benches/performance.rs:34-61,
// examples/dot_cube.rs:200-231)
// ---------------------------------------------------------------------------------------------
Ray make_pixel_ray(const Camera& cam, uint32_t w, uint32_t h, uint32_t x, uint32_t y) {
    const V3f up = {0.0f, 1.0f, 0.0f};
    const V3f right = normalized(cross(up, cam.direction));
    const float pixel_width = cam.glass_width / (float)w;
    const float pixel_height = cam.glass_height / (float)h;
    const V3f bottom_left = cam.origin + (cam.direction * cam.glass_distance) - (up * (cam.glass_height / 2.0f)) -
                            (right * (cam.glass_width / 2.0f));
    const V3f glass_point = bottom_left + right * (float)x * pixel_width + up * (float)y * pixel_height;
    return Ray{cam.origin, normalized(glass_point - cam.origin)};
}

// examples/cpu_render.rs:119-136
uint32_t shade_pixel(const Hit& hit, V3f l) {
    auto as_u8 = [](float v) -> uint32_t {  // `f32 as u8`: truncating, saturating, NaN -> 0
        if (!(v > 0.0f)) return 0u;
        return v >= 255.0f ? 255u : (uint32_t)v;
    };
    if (!hit.hit) return 0xFF808080u;  // Rgb([128, 128, 128])
    if (hit.entry.kind != EntryKind::Visual && hit.entry.kind != EntryKind::Complex) return 0xFF000000u;
    // normal.dot(&light) (vector.rs:182-184); both normalized, so the strength is in 0..1
    const float dot = hit.normal.x * l.x + hit.normal.y * l.y + hit.normal.z * l.z;
    const float diffuse_light_strength = 1.0f - (dot / 2.0f + 0.5f);
    const Albedo a = hit.entry.albedo;
    return as_u8((float)a.r * diffuse_light_strength) | (as_u8((float)a.g * diffuse_light_strength) << 8) |
           (as_u8((float)a.b * diffuse_light_strength) << 16) | 0xFF000000u;
}

}  // namespace svxo
