// MagicaVoxel `.vox` import: Octree::load_vox_file / MIPMapStrategy::load_vox_file of the reference
// (src/convert/magicavoxel.rs), the input format of its large examples (minecraft.rs:57-60, sponza.rs:66-67).
//
// The reference sits on the third-party `dot_vox` 5.1.1 parser (Cargo.toml), which is not in the checkout: the chunk reader
// below follows the published MagicaVoxel format (MAIN { SIZE XYZI ... nTRN nGRP nSHP ... RGBA }) and dot_vox's conventions
// that decide results - the colour index stored in XYZI is 1-based and used minus one (saturating), the RGBA chunk's 256
// entries are the palette in file order. Everything after parsing restates magicavoxel.rs line by line (cited below).
// Two inputs the reference can load and this loader refuses, instead of inventing data:
//   * a file without an RGBA chunk - dot_vox substitutes MagicaVoxel's built-in default palette, which is not reproduced here;
//   * a file without a scene graph - the reference panics on `vox_tree.scenes[0]` (magicavoxel.rs:112).
// Parity status: the parsing layer is unpinned (no dot_vox, no Rust here); the placement arithmetic is pinned by the
// reference's rotation KAT (:392-413) and cross-checked against the independent Python reader of round 1 on the reference's
// own assets (tests/test_vox_import.py). `file:line` citations are relative to the reference checkout.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "vox_import.hpp"

namespace svx {

namespace {

struct Reader {
    const uint8_t* p;
    size_t n, pos = 0;
    bool ok = true;
    bool need(size_t k) {
        if (!ok || k > n - pos) ok = false;
        return ok;
    }
    int32_t i32() {
        if (!need(4)) return 0;
        int32_t v;
        std::memcpy(&v, p + pos, 4);
        pos += 4;
        return v;
    }
    std::string str() {
        const int32_t len = i32();
        if (len < 0 || !need((size_t)len)) {
            ok = false;
            return {};
        }
        std::string s(reinterpret_cast<const char*>(p + pos), (size_t)len);
        pos += (size_t)len;
        return s;
    }
    std::map<std::string, std::string> dict() {
        std::map<std::string, std::string> d;
        const int32_t k = i32();
        for (int32_t i = 0; ok && i < k; ++i) {
            std::string key = str();
            d[key] = str();
        }
        return d;
    }
};

struct Mat3 {  // integer 3x3, row major; rotations of a .vox scene graph only hold 0 / +-1
    int m[3][3];
};
const Mat3 IDENTITY{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
Mat3 mul(const Mat3& a, const Mat3& b) {
    Mat3 r{};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
struct Vec3i {
    int64_t x, y, z;
};
// V3c::transformed, magicavoxel.rs:92-100
Vec3i transformed(const Vec3i& v, const Mat3& m) {
    return {v.x * m.m[0][0] + v.y * m.m[0][1] + v.z * m.m[0][2], v.x * m.m[1][0] + v.y * m.m[1][1] + v.z * m.m[1][2],
            v.x * m.m[2][0] + v.y * m.m[2][1] + v.z * m.m[2][2]};
}
// convert_coordinate between Rzup and Lyup swaps y and z either way (src/spatial/math/mod.rs:164-201)
Vec3i swap_yz(const Vec3i& v) { return {v.x, v.z, v.y}; }
int64_t half(int64_t v) { return v / 2; }  // i32 `/ 2` truncates toward zero, like C++

enum NodeType { TRANSFORM, GROUP, SHAPE };
struct SceneNode {
    NodeType type;
    int32_t child = -1;                                           // transform
    std::vector<std::map<std::string, std::string>> frames;       // transform
    std::vector<int32_t> children;                                // group
    std::vector<std::pair<int32_t, std::map<std::string, std::string>>> models;  // shape: (model id, attributes)
};
struct Model {
    int32_t sx = 0, sy = 0, sz = 0;
    std::vector<uint8_t> xyzi;  // 4 bytes per voxel: x, y, z, colour index (already minus one)
};
struct VoxFile {
    std::vector<Model> models;
    std::vector<SceneNode> scenes;  // indexed by node id, as dot_vox does (`vox_tree.scenes[id]`)
    std::vector<svx_albedo> palette;
};

// `str::parse::<i32>()`: an optional sign, then digits only (no white space), and the value must fit an i32
bool parse_int(const std::string& s, long* out) {
    size_t i = (!s.empty() && (s[0] == '-' || s[0] == '+')) ? 1 : 0;
    if (i == s.size()) return false;
    int64_t v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
        if (v > (int64_t)INT32_MAX + 1) return false;
    }
    if (s[0] == '-') v = -v;
    if (v > INT32_MAX) return false;
    *out = (long)v;
    return true;
}

int32_t parse(const uint8_t* data, size_t len, VoxFile* out, std::string* why) {
    if (len < 20 || std::memcmp(data, "VOX ", 4) != 0 || std::memcmp(data + 8, "MAIN", 4) != 0) {
        *why = "not a MagicaVoxel file";
        return SVX_E_DECODE;
    }
    Reader r{data, len};
    r.pos = 12;
    const int32_t main_content = r.i32(), main_children = r.i32();
    if (main_content < 0 || main_children < 0 || !r.need((size_t)main_content)) {
        *why = "malformed MAIN chunk";
        return SVX_E_DECODE;
    }
    r.pos += (size_t)main_content;
    const size_t end = std::min(len, r.pos + (size_t)main_children);
    bool have_size = false;
    Model pending;
    while (r.pos + 12 <= end) {
        char id[5] = {0, 0, 0, 0, 0};
        std::memcpy(id, data + r.pos, 4);
        r.pos += 4;
        const int32_t n_content = r.i32(), n_children = r.i32();
        if (n_content < 0 || n_children < 0 || !r.need((size_t)n_content)) {
            *why = "malformed chunk header";
            return SVX_E_DECODE;
        }
        const size_t body = r.pos, next = body + (size_t)n_content + (size_t)n_children;
        Reader c{data, body + (size_t)n_content};
        c.pos = body;
        if (!std::strcmp(id, "SIZE")) {
            pending = Model();
            pending.sx = c.i32();
            pending.sy = c.i32();
            pending.sz = c.i32();
            have_size = c.ok;
        } else if (!std::strcmp(id, "XYZI")) {
            const int32_t n = c.i32();
            if (n < 0 || !c.need((size_t)n * 4) || !have_size) {
                *why = "malformed XYZI chunk";
                return SVX_E_DECODE;
            }
            pending.xyzi.assign(data + c.pos, data + c.pos + (size_t)n * 4);
            for (size_t i = 3; i < pending.xyzi.size(); i += 4)  // dot_vox: "i is 1 less than the value stored in the source file"
                pending.xyzi[i] = pending.xyzi[i] ? (uint8_t)(pending.xyzi[i] - 1) : 0;
            out->models.push_back(std::move(pending));
            pending = Model();
            have_size = false;
        } else if (!std::strcmp(id, "RGBA")) {
            if (!c.need(1024)) {
                *why = "malformed RGBA chunk";
                return SVX_E_DECODE;
            }
            out->palette.resize(256);
            for (int i = 0; i < 256; ++i) out->palette[i] = svx_albedo{data[c.pos + 4 * i], data[c.pos + 4 * i + 1], data[c.pos + 4 * i + 2], data[c.pos + 4 * i + 3]};
        } else if (!std::strcmp(id, "nTRN") || !std::strcmp(id, "nGRP") || !std::strcmp(id, "nSHP")) {
            const int32_t node_id = c.i32();
            c.dict();  // node attributes (name, hidden): not used by the loader
            SceneNode node;
            if (id[1] == 'T') {
                node.type = TRANSFORM;
                node.child = c.i32();
                c.i32();  // reserved
                c.i32();  // layer
                const int32_t n_frames = c.i32();
                for (int32_t f = 0; c.ok && f < n_frames; ++f) node.frames.push_back(c.dict());
            } else if (id[1] == 'G') {
                node.type = GROUP;
                const int32_t n = c.i32();
                for (int32_t k = 0; c.ok && k < n; ++k) node.children.push_back(c.i32());
            } else {
                node.type = SHAPE;
                const int32_t n = c.i32();
                for (int32_t k = 0; c.ok && k < n; ++k) {
                    const int32_t model_id = c.i32();
                    node.models.emplace_back(model_id, c.dict());
                }
            }
            if (!c.ok || node_id < 0 || node_id > (1 << 24)) {
                *why = "malformed scene graph chunk";
                return SVX_E_DECODE;
            }
            if ((size_t)node_id >= out->scenes.size()) out->scenes.resize((size_t)node_id + 1, SceneNode{GROUP});
            out->scenes[(size_t)node_id] = std::move(node);
        }
        if (next > end || next < body) break;
        r.pos = next;
    }
    if (out->models.empty()) {
        *why = "no models in the file";
        return SVX_E_DECODE;
    }
    if (out->palette.empty()) {
        *why = "no RGBA chunk: MagicaVoxel's built-in default palette (dot_vox substitutes it) is not reproduced by this loader";
        return SVX_E_DECODE;
    }
    if (out->scenes.empty() || out->scenes[0].type != TRANSFORM) {
        *why = "no scene graph with a transform root (the reference panics here, magicavoxel.rs:112-127)";
        return SVX_E_DECODE;
    }
    return SVX_OK;
}

// parse_rotation_matrix, magicavoxel.rs:60-90: bits 0-1 / 2-3 give the column of the non-zero entry of rows 0 / 1 (the
// third row takes the remaining column), bits 4-6 the signs
Mat3 rotation_from_byte(uint8_t b) {
    Mat3 m{};
    const int c0 = b & 3, c1 = (b >> 2) & 3, c2 = (~(c0 ^ c1)) & 3;
    if (c0 < 3) m.m[0][c0] = (b & 0x10) ? -1 : 1;
    if (c1 < 3) m.m[1][c1] = (b & 0x20) ? -1 : 1;
    if (c2 < 3) m.m[2][c2] = (b & 0x40) ? -1 : 1;
    return m;
}

struct Placed {
    const Model* model;
    Vec3i translation;
    Mat3 rotation;
};

// iterate_vox_tree, magicavoxel.rs:105-197, frame 0: an explicit stack of (node, translation, rotation, child index)
int32_t place_models(const VoxFile& f, std::vector<Placed>* out, std::string* why) {
    struct Item {
        int32_t node;
        Vec3i t;
        Mat3 r;
        uint32_t index;
    };
    const size_t frame = 0;
    std::vector<Item> stack;
    stack.push_back({f.scenes[0].child, {0, 0, 0}, IDENTITY, 0});
    size_t guard = 0;
    while (!stack.empty()) {
        if (++guard > (size_t)1 << 24 || stack.size() > 4096) {
            *why = "scene graph does not terminate";
            return SVX_E_DECODE;
        }
        const Item top = stack.back();
        if (top.node < 0 || (size_t)top.node >= f.scenes.size()) {
            *why = "scene graph refers to a missing node";
            return SVX_E_DECODE;
        }
        const SceneNode& node = f.scenes[(size_t)top.node];
        if (node.type == TRANSFORM) {
            if (node.frames.empty()) {
                *why = "transform node without frames";
                return SVX_E_DECODE;
            }
            const auto& fr = node.frames[frame < node.frames.size() ? frame : 0];
            Vec3i t = top.t;
            auto it = fr.find("_t");
            if (it != fr.end()) {
                // `translation + t.split(" ").map(|x| x.parse().expect(..)).collect::<Vec<i32>>().into()` (:136-141): EVERY
                // part must be an integer, the first three are used (vector.rs:366-371)
                long v[3] = {0, 0, 0};
                const std::string& s = it->second;
                size_t a = 0;
                int k = 0;
                for (;;) {
                    const size_t b = std::min(s.find(' ', a), s.size());
                    long part = 0;
                    if (!parse_int(s.substr(a, b - a), &part)) {
                        *why = "translation: a part is not an integer";
                        return SVX_E_DECODE;
                    }
                    if (k < 3) v[k] = part;
                    ++k;
                    if (b == s.size()) break;
                    a = b + 1;
                }
                if (k < 3) {
                    *why = "translation has fewer than three parts";
                    return SVX_E_DECODE;
                }
                t = {t.x + v[0], t.y + v[1], t.z + v[2]};
            }
            // a transform WITHOUT `_r` resets the orientation to the identity (:148-157), it does not inherit
            Mat3 rot = IDENTITY;
            it = fr.find("_r");
            if (it != fr.end()) {
                long b = 0;
                if (!parse_int(it->second, &b) || b < 0 || b > 255) {
                    *why = "rotation is not a byte";
                    return SVX_E_DECODE;
                }
                rot = mul(top.r, rotation_from_byte((uint8_t)b));
            }
            if (top.index == 0) {
                stack.back().index += 1;
                stack.push_back({node.child, t, rot, 0});
            } else {
                stack.pop_back();
            }
        } else if (node.type == GROUP) {
            if (top.index < node.children.size()) {
                stack.back().index += 1;
                stack.push_back({node.children[top.index], top.t, top.r, 0});
            } else {
                stack.pop_back();
            }
        } else {
            for (const auto& m : node.models) {
                long fno = 0;
                auto it = m.second.find("_f");
                if (it != m.second.end() && !parse_int(it->second, &fno)) {
                    *why = "model frame attribute is not an integer";
                    return SVX_E_DECODE;
                }
                if ((size_t)fno != frame) continue;
                if (m.first < 0 || (size_t)m.first >= f.models.size()) {
                    *why = "shape refers to a missing model";
                    return SVX_E_DECODE;
                }
                out->push_back({&f.models[(size_t)m.first], top.t, top.r});
            }
            stack.pop_back();
            if (!stack.empty()) stack.back().index += 1;
        }
    }
    return SVX_OK;
}

struct Layout {
    VoxFile file;
    std::vector<Placed> placed;
    Vec3i min_lyup{0, 0, 0};
    uint32_t tree_size = 0;
};

// load_vox_file_internal (:297-347) + the tree size of load_vox_file (:266-271)
int32_t layout_of(const uint8_t* data, size_t len, Layout* L, std::string* why) {
    int32_t s = parse(data, len, &L->file, why);
    if (s == SVX_OK) s = place_models(L->file, &L->placed, why);
    if (s != SVX_OK) return s;
    Vec3i lo{INT32_MAX, INT32_MAX, INT32_MAX}, hi{INT32_MIN, INT32_MIN, INT32_MIN};
    for (const Placed& p : L->placed) {
        const Vec3i h = transformed({p.model->sx, p.model->sy, p.model->sz}, p.rotation);
        const Vec3i hh{half(h.x), half(h.y), half(h.z)};
        lo = {std::min({lo.x, p.translation.x - hh.x, p.translation.x + hh.x}), std::min({lo.y, p.translation.y - hh.y, p.translation.y + hh.y}),
              std::min({lo.z, p.translation.z - hh.z, p.translation.z + hh.z})};
        hi = {std::max({hi.x, p.translation.x - hh.x, p.translation.x + hh.x}), std::max({hi.y, p.translation.y - hh.y, p.translation.y + hh.y}),
              std::max({hi.z, p.translation.z - hh.z, p.translation.z + hh.z})};
    }
    if (L->placed.empty()) {
        *why = "the scene graph places no model in frame 0";
        return SVX_E_DECODE;
    }
    L->min_lyup = swap_yz(lo);
    const Vec3i ext = swap_yz({hi.x - lo.x, hi.y - lo.y, hi.z - lo.z});
    const int64_t extent = std::max({ext.x, ext.y, ext.z});
    if (extent > (1 << 30)) {
        *why = "scene extent beyond 2^30";
        return SVX_E_DECODE;
    }
    // `(tree_size as f32).log2().ceil() as u32` then `2_u32.pow(..)` (:268-270); an extent of 0 gives -inf -> 0 -> size 1
    const float l2 = std::ceil(std::log2((float)extent));
    const uint32_t exponent = l2 > 0.0f ? (uint32_t)l2 : 0u;
    L->tree_size = 1u << exponent;
    return SVX_OK;
}

// load_vox_data_internal, :349-385: every voxel of every placed model, in file order, as Octree::insert of a Visual entry
int32_t insert_models(const Layout& L, HostOctree* tree, std::string* why) {
    const Vec3i min_rzup = swap_yz(L.min_lyup);
    for (const Placed& p : L.placed) {
        const Vec3i h = transformed({p.model->sx, p.model->sy, p.model->sz}, p.rotation);
        const Vec3i hh{half(h.x), half(h.y), half(h.z)};
        // "if the index delta is negative (because of orientation) ... a correction in every dimension where the index is below 0"
        const Vec3i bottom_left{p.translation.x - hh.x - min_rzup.x + (hh.x < 0 ? -1 : 0), p.translation.y - hh.y - min_rzup.y + (hh.y < 0 ? -1 : 0),
                                p.translation.z - hh.z - min_rzup.z + (hh.z < 0 ? -1 : 0)};
        const std::vector<uint8_t>& v = p.model->xyzi;
        for (size_t i = 0; i + 3 < v.size(); i += 4) {
            const Vec3i rot = transformed({v[i], v[i + 1], v[i + 2]}, p.rotation);
            const Vec3i pos = swap_yz({bottom_left.x + rot.x, bottom_left.y + rot.y, bottom_left.z + rot.z});
            if (pos.x < 0 || pos.y < 0 || pos.z < 0 || pos.x >= tree->size() || pos.y >= tree->size() || pos.z >= tree->size()) {
                *why = "a voxel lies outside the tree (the reference panics: \"inserting into octree at at invalid position\", :371-376)";
                return SVX_E_INVALID_POSITION;
            }
            svx_entry e{};
            e.kind = SVX_ENTRY_VISUAL;
            e.albedo = L.file.palette[v[i + 3]];
            const int32_t s = tree->insert_at_lod_internal(true, (uint32_t)pos.x, (uint32_t)pos.y, (uint32_t)pos.z, 1, e);
            if (s != SVX_OK) {
                *why = "insert failed";
                return s;
            }
        }
    }
    return SVX_OK;
}

}  // namespace

int32_t vox_required_tree_size(const uint8_t* data, size_t len, uint32_t* tree_size, std::string* why) {
    Layout L;
    const int32_t s = layout_of(data, len, &L, why);
    if (s == SVX_OK) *tree_size = L.tree_size;
    return s;
}

int32_t vox_insert_into(const uint8_t* data, size_t len, HostOctree* tree, std::string* why) {
    Layout L;
    const int32_t s = layout_of(data, len, &L, why);
    if (s != SVX_OK) return s;
    if (tree->size() < L.tree_size) {
        *why = "the tree is smaller than the file's extent";
        return SVX_E_INVALID_SIZE;
    }
    return insert_models(L, tree, why);
}

int32_t vox_load(const uint8_t* data, size_t len, uint32_t brick_dim, HostOctree** out, std::string* why) {
    Layout L;
    int32_t s = layout_of(data, len, &L, why);
    if (s != SVX_OK) return s;
    HostOctree* tree = nullptr;
    s = HostOctree::create(L.tree_size, brick_dim, &tree);  // the reference panics when Octree::new refuses (:273-281)
    if (s != SVX_OK) {
        *why = "Octree::new refuses tree size " + std::to_string(L.tree_size) + " with brick dimension " + std::to_string(brick_dim);
        return s;
    }
    s = insert_models(L, tree, why);
    if (s != SVX_OK) {
        delete tree;
        return s;
    }
    *out = tree;
    return SVX_OK;
}

int32_t vox_read_file(const char* path, std::vector<uint8_t>* bytes) {
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return SVX_E_IO;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) {
        std::fclose(f);
        return SVX_E_IO;
    }
    bytes->resize((size_t)n);
    const size_t got = n ? std::fread(bytes->data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n ? SVX_OK : SVX_E_IO;
}

}  // namespace svx
