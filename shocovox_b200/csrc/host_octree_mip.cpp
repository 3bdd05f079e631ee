// MIP maps of the product's host octree: one simplified brick per node, kept up to date by every edit and read by the
// traversal kernel's level-of-detail branch (traverse.cuh: SVX_LOD). Reference: src/octree/mipmap.rs; all `file:line`
// citations are relative to the reference checkout.
//
// A MIP is an ordinary brick of the pooled voxel array (NodeRec::mip), so the device mirrors it, derives its occupancy
// bits and walks it with the same DDA as any leaf brick. Resampling is split in two steps here: GATHER the albedos of
// the sampling range into a flat list (x outer, y, z inner - the reference's visiting order, which fixes the f32
// summation order of the box filter), then REDUCE that list with the level's method.
//
// Determinism: the reference's PointFilter / Posterize keep candidates in a randomly seeded HashMap and pick "the last
// maximum in iteration order", so ties are resolved differently from run to run there. Here (and in the test oracle)
// candidates are kept in first-seen order and a tie goes to the earliest - one of the outcomes the reference can produce.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "host_octree.hpp"

namespace svx {

namespace {

// Rust `f32 as u32` / `as u8` / `as usize`: truncate, saturate, NaN -> 0
inline uint32_t as_u32(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
inline uint8_t as_u8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
// From<V3c<f32>> for V3c<u32>: round, then cast (src/spatial/math/vector.rs:326-336)
inline uint32_t rounded_u32(float v) { return as_u32(std::round(v)); }

inline size_t flat(size_t x, size_t y, size_t z, size_t dim) { return x + y * dim + z * dim * dim; }

inline uint8_t octant_of(float dx, float dy, float dz, float half) {
    return (uint8_t)((dx >= half) + (dz >= half) * 2 + (dy >= half) * 4);
}

inline bool has_albedo(const svx_entry& e) { return e.kind == SVX_ENTRY_VISUAL || e.kind == SVX_ENTRY_COMPLEX; }
inline bool same(const svx_albedo& a, const svx_albedo& b) { return a.r == b.r && a.g == b.g && a.b == b.b && a.a == b.a; }

// Albedo::distance_from, src/octree/detail.rs:82-89
inline float distance(const svx_albedo& a, const svx_albedo& b) {
    const float dr = (float)a.r - (float)b.r, dg = (float)a.g - (float)b.g, db = (float)a.b - (float)b.b,
                da = (float)a.a - (float)b.a;
    return std::sqrt(dr * dr + dg * dg + db * db + da * da);
}

// The posterize buckets: per channel the sum of squares (u32, like the reference's Albedou32) and the member count
struct Bucket {
    uint32_t sum[4];
    uint32_t count;
};
// (sum / count).sqrt() per channel with the reference's two roundings (mipmap.rs:88-98, :47-53)
inline void bucket_colour(const Bucket& b, uint32_t out[4]) {
    for (int c = 0; c < 4; ++c) {
        const uint32_t mean = as_u32(std::round((float)b.sum[c] / (float)b.count));
        out[c] = as_u32(std::round(std::sqrt((float)mean)));
    }
}

inline uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    return h;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// strategy: MIPMapStrategy, mipmap.rs:591-690
// ------------------------------------------------------------------------------------------------------------
void HostOctree::mip_defaults() {
    mips_enabled_ = false;
    mip_methods_.clear();
    mip_methods_[1] = MipSampler{MIP_POINT, 0.0f};
    for (size_t level = 2; level <= 4; ++level) mip_methods_[level] = MipSampler{MIP_BOX, 0.0f};
    mip_thresholds_.clear();
    mip_thresholds_[2] = 0.1f;
    mip_thresholds_[3] = 0.05f;
    mip_thresholds_[4] = 0.02f;
}

void HostOctree::mip_reset() {
    mip_defaults();
    ++revision_;
}

void HostOctree::mip_set_method_at(size_t level, uint32_t method, float thr) {
    MipSampler s{method, 0.0f};
    if (method == MIP_POSTERIZE || method == MIP_POSTERIZE_BD) s.thr = std::min(std::max(thr, 0.0f), 1.0f);
    mip_methods_[level] = s;
}

MipSampler HostOctree::mip_get_method_at(size_t level) const {
    auto it = mip_methods_.find(level);
    return it == mip_methods_.end() ? MipSampler{} : it->second;  // BoxFilter is the default method
}

void HostOctree::mip_set_color_similarity_thr_at(size_t level, float thr) {
    mip_thresholds_[level] = std::min(std::max(thr, 0.0f), 1.0f);
}

float HostOctree::mip_get_color_similarity_at(size_t level) const {
    auto it = mip_thresholds_.find(level);
    return it == mip_thresholds_.end() ? 0.0f : it->second;
}

void HostOctree::switch_albedo_mip_maps(bool enabled) {
    const bool was = mips_enabled_;
    mips_enabled_ = enabled;
    ++revision_;
    if (enabled && !was && nodes_[0].kind != NK_NOTHING) recalculate_mips();
}

void HostOctree::mip_load_strategy(bool enabled, std::map<size_t, MipSampler> methods, std::map<size_t, float> thresholds) {
    mips_enabled_ = enabled;
    mip_methods_ = std::move(methods);
    mip_thresholds_ = std::move(thresholds);
    ++revision_;
}

void HostOctree::mip_load_brick(size_t key, uint8_t kind, uint32_t solid, const uint32_t* voxels) {
    if (key >= nodes_.size()) return;
    NodeRec& n = nodes_[key];
    brick_release(n.mip);
    if (kind == BK_SOLID) {
        n.mip.kind = BK_SOLID;
        n.mip.value = solid;
    } else if (kind == BK_PARTED) {
        const uint32_t h = brick_alloc(NIL);
        std::memcpy(brick_mut(h), voxels, (size_t)vol_ * 4);
        nodes_[key].mip.kind = BK_PARTED;
        nodes_[key].mip.value = h;
    }
}

// ------------------------------------------------------------------------------------------------------------
// reduce: MIPResaplingFunction::execute, mipmap.rs:133-263
// ------------------------------------------------------------------------------------------------------------
bool HostOctree::mip_reduce(const MipSampler& how, const std::vector<svx_albedo>& samples, svx_albedo* out) const {
    if (samples.empty()) return false;
    switch (how.method) {
        default:
        case MIP_BOX: {
            // gamma-2 average: sqrt(mean of squares), f32 sums in visiting order, truncated to u8 (:141-184)
            float acc[4] = {0, 0, 0, 0};
            bool first = true;
            for (const svx_albedo& s : samples) {
                const float sq[4] = {(float)s.r * (float)s.r, (float)s.g * (float)s.g, (float)s.b * (float)s.b,
                                     (float)s.a * (float)s.a};
                for (int c = 0; c < 4; ++c) acc[c] = first ? sq[c] : acc[c] + sq[c];
                first = false;
            }
            const float n = (float)(int32_t)samples.size();
            out->r = as_u8(std::fmin(std::sqrt(acc[0] / n), 255.0f));
            out->g = as_u8(std::fmin(std::sqrt(acc[1] / n), 255.0f));
            out->b = as_u8(std::fmin(std::sqrt(acc[2] / n), 255.0f));
            out->a = as_u8(std::fmin(std::sqrt(acc[3] / n), 255.0f));
            return true;
        }
        case MIP_POINT:
        case MIP_POINT_BD: {
            // the most frequent colour (:185-206)
            struct Tally {
                svx_albedo colour;
                uint32_t count;
            };
            std::vector<Tally> tally;
            for (const svx_albedo& s : samples) {
                auto it = std::find_if(tally.begin(), tally.end(), [&](const Tally& t) { return same(t.colour, s); });
                if (it == tally.end())
                    tally.push_back({s, 1});
                else
                    it->count += 1;
            }
            const Tally* best = &tally[0];
            for (const Tally& t : tally)
                if (t.count > best->count) best = &t;
            *out = best->colour;
            return true;
        }
        case MIP_POSTERIZE:
        case MIP_POSTERIZE_BD: {
            // colours within `thr * 255` of a bucket's current colour join it; the biggest bucket wins (:207-261)
            const float limit = how.thr * 255.0f;
            std::vector<Bucket> buckets;
            for (const svx_albedo& s : samples) {
                const uint32_t px[4] = {s.r, s.g, s.b, s.a};
                const uint32_t sq[4] = {px[0] * px[0], px[1] * px[1], px[2] * px[2], px[3] * px[3]};
                size_t home = buckets.size();
                for (size_t i = 0; i < buckets.size() && home == buckets.size(); ++i) {
                    uint32_t colour[4];
                    bucket_colour(buckets[i], colour);
                    // u32 wrap-around like a release build of the reference: (2^32 - d)^2 == d^2 (mod 2^32)
                    uint32_t d2 = 0;
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t d = colour[c] - px[c];
                        d2 += d * d;
                    }
                    if (std::sqrt((float)d2) < limit) home = i;
                }
                // the buckets are keyed by their sums (a HashMap in the reference): writing a key that another bucket
                // already holds replaces that bucket's count; a re-keyed bucket keeps its place in the order
                auto holder_of = [&](const uint32_t sum[4], size_t except) {
                    for (size_t i = 0; i < buckets.size(); ++i)
                        if (i != except && std::memcmp(buckets[i].sum, sum, 4 * sizeof(uint32_t)) == 0) return i;
                    return buckets.size();
                };
                if (home < buckets.size()) {
                    Bucket& b = buckets[home];
                    for (int c = 0; c < 4; ++c) b.sum[c] += sq[c];
                    b.count += 1;
                    const size_t other = holder_of(b.sum, home);
                    if (other < buckets.size()) buckets.erase(buckets.begin() + (ptrdiff_t)other);
                } else {
                    const size_t other = holder_of(sq, buckets.size());
                    if (other < buckets.size())
                        buckets[other].count = 1;
                    else
                        buckets.push_back(Bucket{{sq[0], sq[1], sq[2], sq[3]}, 1});
                }
            }
            const Bucket* best = &buckets[0];
            for (const Bucket& b : buckets)
                if (b.count > best->count) best = &b;
            uint32_t colour[4];
            bucket_colour(*best, colour);
            out->r = (uint8_t)std::min(colour[0], 255u);
            out->g = (uint8_t)std::min(colour[1], 255u);
            out->b = (uint8_t)std::min(colour[2], 255u);
            out->a = (uint8_t)std::min(colour[3], 255u);
            return true;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// gather from the children's MIPs (Internal nodes, mipmap.rs:458-511). `start` is in the node's 2*dim MIP space.
// Returns false when the child under `start` does not exist (the whole cell is then skipped, :460-465).
// ------------------------------------------------------------------------------------------------------------
bool HostOctree::mip_gather_children(size_t key, const uint32_t start[3], std::vector<svx_albedo>* out) const {
    const NodeRec& n = nodes_[key];
    const float fdim = (float)dim_;
    auto child_at = [&](uint32_t x, uint32_t y, uint32_t z, uint8_t* oct) -> uint32_t {
        *oct = octant_of((float)x, (float)y, (float)z, fdim);
        return n.link == LK_CHILDREN ? n.child[*oct] : NIL;
    };
    uint8_t oct;
    if (child_at(start[0], start[1], start[2], &oct) == NIL) return false;
    for (uint32_t x = start[0]; x < start[0] + 2; ++x)
        for (uint32_t y = start[1]; y < start[1] + 2; ++y)
            for (uint32_t z = start[2]; z < start[2] + 2; ++z) {
                // brick_dim == 1: the 2-cell straddles children, so the octant is re-derived per sample (:469-476)
                const uint32_t child = child_at(x, y, z, &oct);
                if (child == NIL || child >= nodes_.size()) continue;
                const BrickRef& m = nodes_[child].mip;
                if (m.kind == BK_EMPTY) continue;
                uint32_t value = m.value;
                if (m.kind == BK_PARTED) {
                    const uint32_t lx = x - (oct & 1) * dim_, ly = y - ((oct >> 2) & 1) * dim_, lz = z - ((oct >> 1) & 1) * dim_;
                    value = brick_data(m.value)[flat(lx, ly, lz, dim_)];
                }
                const svx_entry e = resolve(value);
                if (has_albedo(e)) out->push_back(e.albedo);
            }
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// update_mip, mipmap.rs:296-584: refresh the one MIP voxel of node `key` that covers position (x, y, z)
// ------------------------------------------------------------------------------------------------------------
void HostOctree::update_mip(size_t key, const BoundsF& nb, uint32_t x, uint32_t y, uint32_t z) {
    if (!mips_enabled_) return;
    NodeRec& node = nodes_[key];
    if (node.kind == NK_NOTHING) return;
    if (node.kind == NK_UNIFORM) {  // its content is its own MIP (:331-337)
        brick_release(node.mip);
        return;
    }
    const size_t level = (size_t)as_u32(std::log2(nb.size / (float)dim_));
    const MipSampler how = mip_get_method_at(level);
    const uint32_t node_size = as_u32(nb.size);
    const uint32_t pos[3] = {x, y, z};

    std::vector<svx_albedo>& samples = mip_scratch_;
    samples.clear();
    const bool from_voxels = node.kind == NK_LEAF || how.method == MIP_POINT_BD;  // :309-314 names PointFilterBD only
    if (from_voxels) {
        // sample the voxels themselves through get(): Leaf -> the 2^3 voxels under the MIP voxel (:338-375);
        // bottom-dominant Internal -> the whole (size/dim)^3 region (:376-411)
        // insert_at_lod with a size <= brick_dim can split leaves into nodes smaller than a brick (insert.rs:137-147):
        // there `size / dim` is 0 and the reference panics on `position % 0`; such a node keeps its MIP as it is
        uint32_t span, start[3];
        if (node_size / dim_ == 0) return;
        if (node.kind == NK_LEAF) {
            span = std::min(node_size / dim_, dim_ * 2);
            for (int c = 0; c < 3; ++c) {
                const uint32_t scaled = (pos[c] - pos[c] % span) * 2 * dim_;
                start[c] = rounded_u32(std::floor((float)scaled / nb.size));
            }
        } else {
            span = node_size / dim_;
            for (int c = 0; c < 3; ++c) start[c] = rounded_u32(std::floor((float)(pos[c] - pos[c] % span)));
        }
        for (uint32_t sx = start[0]; sx < start[0] + span; ++sx)
            for (uint32_t sy = start[1]; sy < start[1] + span; ++sy)
                for (uint32_t sz = start[2]; sz < start[2] + span; ++sz) {
                    const svx_entry e = get_from(key, nb, sx, sy, sz);
                    if (has_albedo(e)) samples.push_back(e.albedo);
                }
    } else {
        // Internal: the 2^3 voxels of the children's MIPs under this MIP voxel, addressed in 2*dim space (:412-446)
        const float origin[3] = {nb.x, nb.y, nb.z};
        uint32_t start[3];
        for (int c = 0; c < 3; ++c) {
            const float in_mip_space = ((float)pos[c] - origin[c]) * 2.0f * (float)dim_ / nb.size;
            start[c] = rounded_u32(std::floor(in_mip_space));
            start[c] -= start[c] % 2;
        }
        if (!mip_gather_children(key, start, &samples)) return;
    }

    svx_albedo colour;
    if (!mip_reduce(how, samples, &colour)) return;  // nothing sampled: the MIP voxel keeps what it had (:513, :550)

    // palette entry: reuse the first colour within the level's similarity threshold, else add a new one (:513-548)
    svx_entry visual{};
    visual.kind = SVX_ENTRY_VISUAL;
    visual.albedo = colour;
    uint32_t value = NIL;
    bool reused = false;
    auto thr = mip_thresholds_.find(level);
    if (thr != mip_thresholds_.end()) {
        const float limit = thr->second * 255.0f;
        for (size_t i = 0; i < colors_.size() && !reused; ++i)
            if (distance(colour, colors_[i]) < limit) {
                value = ((uint32_t)i & 0xFFFFu) | 0xFFFF0000u;  // pix_visual, node.rs:354
                reused = true;
            }
    }
    if (!reused) value = add_to_palette(visual);

    // write it (:550-581); the index is matrix_index_for(node bounds, position, dim), math/mod.rs:44-77
    size_t cell[3];
    const float origin[3] = {nb.x, nb.y, nb.z};
    for (int c = 0; c < 3; ++c) {
        const float v = std::floor(((float)pos[c] - origin[c]) * (float)dim_ / nb.size);
        cell[c] = (size_t)as_u32(std::round(v));
    }
    const size_t index = flat(cell[0], cell[1], cell[2], dim_);
    if (index >= vol_) return;
    if (nodes_[key].mip.kind != BK_PARTED) {
        const uint32_t fill = nodes_[key].mip.kind == BK_SOLID ? nodes_[key].mip.value : NIL;
        const uint32_t h = brick_alloc(fill);
        nodes_[key].mip.kind = BK_PARTED;
        nodes_[key].mip.value = h;
    }
    brick_mut(nodes_[key].mip.value)[index] = value;
}

// recalculate_mip, mipmap.rs:875-892
void HostOctree::recalculate_mip(size_t key, const BoundsF& nb) {
    if (!mips_enabled_) return;
    for (uint32_t x = 0; x < dim_; ++x)
        for (uint32_t y = 0; y < dim_; ++y)
            for (uint32_t z = 0; z < dim_; ++z) {
                // min + round((x, y, z) * size / dim), then V3c<u32>::from rounds again
                const float px = nb.x + std::round((float)x * nb.size / (float)dim_);
                const float py = nb.y + std::round((float)y * nb.size / (float)dim_);
                const float pz = nb.z + std::round((float)z * nb.size / (float)dim_);
                update_mip(key, nb, rounded_u32(px), rounded_u32(py), rounded_u32(pz));
            }
}

// recalculate_mips, mipmap.rs:798-855: drop every MIP, then rebuild bottom-up (children in octant order first)
void HostOctree::recalculate_mips() {
    ++revision_;
    for (NodeRec& n : nodes_) brick_release(n.mip);
    struct Frame {
        uint32_t key;
        BoundsF b;
        uint8_t next;
    };
    std::vector<Frame> stack;
    stack.push_back({0u, BoundsF{0, 0, 0, (float)size_}, 0});
    while (!stack.empty()) {
        Frame& f = stack.back();
        const NodeRec& n = nodes_[f.key];
        if (n.kind == NK_NOTHING) {  // only an empty root gets here (the reference panics: unreachable!())
            stack.pop_back();
            continue;
        }
        if (n.kind == NK_INTERNAL && f.next < 8) {
            const uint8_t o = f.next++;
            const uint32_t c = n.link == LK_CHILDREN ? n.child[o] : NIL;
            if (key_is_valid(c) && nodes_[c].kind != NK_NOTHING) {
                const float half = f.b.size / 2.0f;
                const BoundsF cb{f.b.x + (float)(o & 1) * half, f.b.y + (float)((o >> 2) & 1) * half,
                                 f.b.z + (float)((o >> 1) & 1) * half, half};
                stack.push_back({c, cb, 0});
            }
            continue;
        }
        const Frame done = f;
        stack.pop_back();
        recalculate_mip(done.key, done.b);
    }
}

// sample_root_mip, mipmap.rs:897-937 (the reference's test hook)
svx_entry HostOctree::sample_root_mip(uint32_t octant, uint32_t x, uint32_t y, uint32_t z) const {
    svx_entry none{};
    size_t key = 0;
    if (octant != 8) {
        const NodeRec& root = nodes_[0];
        key = (root.link == LK_CHILDREN && octant < 8) ? root.child[octant] : NIL;
    }
    if (!key_is_valid(key)) return none;
    const BrickRef& m = nodes_[key].mip;
    if (m.kind == BK_EMPTY) return none;
    if (m.kind == BK_SOLID) return resolve(m.value);
    if (x >= dim_ || y >= dim_ || z >= dim_) return none;
    return resolve(brick_data(m.value)[flat(x, y, z, dim_)]);
}

// digest of the strategy and of the MIP brick of every reachable node (same definition as the test oracle's)
uint64_t HostOctree::mip_hash_node(size_t key) const {
    const NodeRec& n = nodes_[key];
    uint64_t h = mix(0x313D, hash_brick(n.mip));
    if (n.kind == NK_INTERNAL)
        for (int o = 0; o < 8; ++o) {
            const uint32_t c = n.link == LK_CHILDREN ? n.child[o] : NIL;
            h = mix(h, key_is_valid(c) ? mip_hash_node(c) : 0x5EED);
        }
    return h;
}

uint64_t HostOctree::mip_hash() const {
    uint64_t h = mix(0x57A7, mips_enabled_ ? 1 : 0);
    for (const auto& m : mip_methods_) {
        uint32_t bits;
        std::memcpy(&bits, &m.second.thr, 4);
        h = mix(mix(mix(h, m.first), m.second.method), bits);
    }
    h = mix(h, 0x7447);
    for (const auto& m : mip_thresholds_) {
        uint32_t bits;
        std::memcpy(&bits, &m.second, 4);
        h = mix(mix(h, m.first), bits);
    }
    return mix(h, mip_hash_node(0));
}

}  // namespace svx
