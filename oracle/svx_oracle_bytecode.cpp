// ORACLE — TEST INFRASTRUCTURE ONLY (see svx_oracle.hpp).
//
// Restatement of the reference's bencode persistence for Octree<u32>: src/convert/bytecode.rs (Albedo :11-62,
// BrickData :64-160, NodeContent :162-287, NodeChildren :289-340, MIPMapStrategy :342-433, MIPResamplingMethods
// :435-484, Octree :575-673) and src/object_pool.rs:25-137 (ReusableItem, ObjectPool), over the oracle's own enum-shaped
// data model. Decoding goes through a generic bencode document tree (the way `bendy` hands out Objects), encoding
// appends to a string. Parity UNPINNED against Rust output: the reference's tests (src/convert/bytecode_tests.rs) are
// round trips without literal bytes, and Rust cannot run here; tests/test_bytecode.py pins the format with a
// hand-assembled literal instead and cross-checks this restatement against the product's independent codec.
#include <cstring>
#include <memory>
#include <string>

#include "svx_oracle.hpp"

namespace svxo {

namespace {

// ---- bencode document ------------------------------------------------------------------------------------------
struct Item {
    enum Kind { Int, Bytes, List } kind = Int;
    uint64_t integer = 0;
    std::string bytes;
    std::vector<Item> list;
};

bool parse_item(const uint8_t*& p, const uint8_t* end, Item* out, int depth = 0) {
    if (p >= end || depth > 64) return false;
    if (*p == 'i') {
        ++p;
        out->kind = Item::Int;
        uint64_t v = 0;
        bool any = false;
        while (p < end && *p >= '0' && *p <= '9') {
            v = v * 10 + (uint64_t)(*p - '0');
            ++p;
            any = true;
        }
        if (!any || p >= end || *p != 'e') return false;
        ++p;
        out->integer = v;
        return true;
    }
    if (*p == 'l') {
        ++p;
        out->kind = Item::List;
        while (p < end && *p != 'e') {
            out->list.emplace_back();
            if (!parse_item(p, end, &out->list.back(), depth + 1)) return false;
        }
        if (p >= end) return false;
        ++p;
        return true;
    }
    if (*p >= '0' && *p <= '9') {
        size_t n = 0;
        while (p < end && *p >= '0' && *p <= '9') {
            n = n * 10 + (size_t)(*p - '0');
            ++p;
        }
        if (p >= end || *p != ':' || (size_t)(end - p - 1) < n) return false;
        ++p;
        out->kind = Item::Bytes;
        out->bytes.assign((const char*)p, n);
        p += n;
        return true;
    }
    return false;
}

void put_int(std::string& s, uint64_t v) { s += "i" + std::to_string(v) + "e"; }
void put_str(std::string& s, const char* lit) { s += std::to_string(std::strlen(lit)) + ":" + lit; }

// BrickData<T>::encode, bytecode.rs:70-90
void put_brick(std::string& s, const Brick& b) {
    switch (b.kind) {
        case BrickKind::Empty: put_str(s, "#b"); break;
        case BrickKind::Solid:
            s += "l";
            put_str(s, "#b#");
            put_int(s, b.solid);
            s += "e";
            break;
        case BrickKind::Parted:
            s += "l";
            put_str(s, "##b#");
            put_int(s, b.data.size());
            for (uint32_t v : b.data) put_int(s, v);
            put_str(s, "#");
            s += "e";
            break;
    }
}

// BrickData<T>::decode_bencode_object, bytecode.rs:97-159
bool get_brick(const Item& it, Brick* b) {
    *b = Brick();
    if (it.kind == Item::Bytes) return it.bytes == "#b";
    if (it.kind != Item::List || it.list.empty() || it.list[0].kind != Item::Bytes) return false;
    if (it.list[0].bytes == "#b#") {
        if (it.list.size() < 2 || it.list[1].kind != Item::Int) return false;
        b->kind = BrickKind::Solid;
        b->solid = (uint32_t)it.list[1].integer;
        return true;
    }
    if (it.list[0].bytes == "##b#") {
        if (it.list.size() < 2 || it.list[1].kind != Item::Int) return false;
        const size_t len = (size_t)it.list[1].integer;
        if (it.list.size() < 2 + len) return false;
        b->kind = BrickKind::Parted;
        b->data.resize(len);
        for (size_t i = 0; i < len; ++i) {
            if (it.list[2 + i].kind != Item::Int) return false;
            b->data[i] = (uint32_t)it.list[2 + i].integer;
        }
        return true;
    }
    return false;
}

}  // namespace

// Octree<T>::encode, bytecode.rs:579-598
std::string octree_to_bytes(const Octree& t) {
    std::string s;
    s += "l";
    put_int(s, t.auto_simplify ? 1 : 0);
    put_int(s, t.octree_size);
    put_int(s, t.brick_dim);
    // ObjectPool::encode, object_pool.rs:102-107 ; ReusableItem::encode :30-35 ; NodeContent::encode bytecode.rs:168-195
    s += "l";
    put_int(s, t.nodes.first_available);
    s += "l";
    for (size_t i = 0; i < t.nodes.len(); ++i) {
        const Node& n = t.nodes.item[i];
        s += "l";
        put_int(s, t.nodes.reserved[i] ? 1 : 0);
        switch (n.kind) {
            case NodeKind::Nothing: put_str(s, "#"); break;
            case NodeKind::Internal:
                s += "l";
                put_str(s, "##");
                put_int(s, n.occupied_bits);
                s += "e";
                break;
            case NodeKind::Leaf:
                s += "l";
                put_str(s, "###");
                for (const Brick& b : n.bricks) put_brick(s, b);
                s += "e";
                break;
            case NodeKind::UniformLeaf:
                s += "l";
                put_str(s, "##u#");
                put_brick(s, n.ubrick);
                s += "e";
                break;
        }
        s += "e";
    }
    s += "ee";
    // Vec<NodeChildren<u32>>, bytecode.rs:291-310
    s += "l";
    for (const Children& c : t.node_children) {
        switch (c.kind) {
            case ChildrenKind::NoChildren: put_str(s, "##x##"); break;
            case ChildrenKind::Children:
                s += "l";
                put_str(s, "##c##");
                for (uint32_t k : c.child) put_int(s, k);
                s += "e";
                break;
            case ChildrenKind::OccupancyBitmap:
                s += "l";
                put_str(s, "##b##");
                put_int(s, c.bitmap);
                s += "e";
                break;
        }
    }
    s += "e";
    // node_mips (bytecode.rs:591): one BrickData per node key; the vector is as long as the node buffer
    // (insert.rs:168-169, detail.rs:356-397 resize it to nodes.len())
    s += "l";
    for (size_t i = 0; i < t.nodes.len(); ++i) put_brick(s, i < t.node_mips.size() ? t.node_mips[i] : Brick());
    s += "e";
    s += "l";
    for (const Albedo& a : t.voxel_color_palette) {  // Albedo::encode, bytecode.rs:13-20
        s += "l";
        put_int(s, a.r);
        put_int(s, a.g);
        put_int(s, a.b);
        put_int(s, a.a);
        s += "e";
    }
    s += "e";
    s += "l";
    for (uint32_t d : t.voxel_data_palette) put_int(s, d);
    s += "e";
    // MIPMapStrategy::encode (bytecode.rs:436-453) with MIPResamplingMethods::encode (:519-535); maps in ascending
    // level order (the reference iterates a HashMap: any order is a valid file)
    auto milli = [](float thr) -> uint64_t {  // `(thr * 1000.) as u32`
        const float v = thr * 1000.0f;
        if (!(v > 0.0f)) return 0;
        return v >= 4294967296.0f ? 0xFFFFFFFFull : (uint64_t)(uint32_t)v;
    };
    s += "l";
    put_int(s, t.mip_map_strategy.enabled ? 1 : 0);
    put_int(s, t.mip_map_strategy.resampling_methods.size());
    for (const auto& m : t.mip_map_strategy.resampling_methods) {
        put_int(s, m.first);
        switch (m.second.method) {
            case MipMethod::BoxFilter: put_int(s, 0); break;
            case MipMethod::PointFilter: put_int(s, 1); break;
            case MipMethod::PointFilterBD: put_int(s, 2); break;
            case MipMethod::Posterize: put_int(s, 3 + milli(m.second.thr)); break;
            case MipMethod::PosterizeBD: put_int(s, 1003 + milli(m.second.thr)); break;
        }
    }
    put_int(s, t.mip_map_strategy.resampling_color_matching_thresholds.size());
    for (const auto& m : t.mip_map_strategy.resampling_color_matching_thresholds) {
        put_int(s, m.first);
        put_int(s, milli(m.second));
    }
    s += "e";
    s += "e";
    return s;
}

// Octree<T>::decode_bencode_object, bytecode.rs:604-672
Status octree_from_bytes(const uint8_t* data, size_t len, Octree** out) {
    *out = nullptr;
    const Status bad = (Status)6;
    Item root;
    const uint8_t* p = data;
    if (!parse_item(p, data + len, &root) || p != data + len) return bad;
    if (root.kind != Item::List || root.list.size() != 9) return bad;
    const std::vector<Item>& f = root.list;
    for (int i = 0; i < 3; ++i)
        if (f[i].kind != Item::Int) return bad;
    if (f[0].integer > 1) return bad;
    Octree* t = nullptr;
    const Status st = Octree::create((uint32_t)f[1].integer, (uint32_t)f[2].integer, &t);
    if (st != OK) return st;
    std::unique_ptr<Octree> hold(t);
    t->auto_simplify = f[0].integer == 1;
    // ObjectPool
    if (f[3].kind != Item::List || f[3].list.size() != 2 || f[3].list[0].kind != Item::Int || f[3].list[1].kind != Item::List)
        return bad;
    t->nodes.item.clear();
    t->nodes.reserved.clear();
    t->nodes.first_available = (size_t)f[3].list[0].integer;
    for (const Item& ri : f[3].list[1].list) {
        if (ri.kind != Item::List || ri.list.size() != 2 || ri.list[0].kind != Item::Int || ri.list[0].integer > 1) return bad;
        const Item& c = ri.list[1];
        Node n;
        if (c.kind == Item::Bytes) {
            if (c.bytes != "#") return bad;
        } else if (c.kind == Item::List && !c.list.empty() && c.list[0].kind == Item::Bytes) {
            const std::string& m = c.list[0].bytes;
            if (m == "##") {
                if (c.list.size() < 2 || c.list[1].kind != Item::Int) return bad;
                n.kind = NodeKind::Internal;
                n.occupied_bits = c.list[1].integer;
            } else if (m == "###") {
                if (c.list.size() < 9) return bad;
                n.kind = NodeKind::Leaf;
                for (int o = 0; o < 8; ++o)
                    if (!get_brick(c.list[1 + o], &n.bricks[o])) return bad;
            } else if (m == "##u#") {
                if (c.list.size() < 2 || !get_brick(c.list[1], &n.ubrick)) return bad;
                n.kind = NodeKind::UniformLeaf;
            } else {
                return bad;
            }
        } else {
            return bad;
        }
        t->nodes.item.push_back(std::move(n));
        t->nodes.reserved.push_back((uint8_t)ri.list[0].integer);
    }
    // node_children
    if (f[4].kind != Item::List) return bad;
    t->node_children.clear();
    for (const Item& c : f[4].list) {
        Children ch;
        if (c.kind == Item::Bytes) {
            if (c.bytes != "##x##") return bad;
        } else if (c.kind == Item::List && !c.list.empty() && c.list[0].kind == Item::Bytes) {
            if (c.list[0].bytes == "##c##") {
                if (c.list.size() < 9) return bad;
                ch.kind = ChildrenKind::Children;
                for (int o = 0; o < 8; ++o) {
                    if (c.list[1 + o].kind != Item::Int) return bad;
                    ch.child[o] = (uint32_t)c.list[1 + o].integer;
                }
            } else if (c.list[0].bytes == "##b##") {
                if (c.list.size() < 2 || c.list[1].kind != Item::Int) return bad;
                ch.kind = ChildrenKind::OccupancyBitmap;
                ch.bitmap = c.list[1].integer;
            } else {
                return bad;
            }
        } else {
            return bad;
        }
        t->node_children.push_back(ch);
    }
    t->node_children.resize(t->nodes.len());
    // node_mips (bytecode.rs:638)
    if (f[5].kind != Item::List) return bad;
    t->node_mips.clear();
    for (const Item& m : f[5].list) {
        Brick b;
        if (!get_brick(m, &b)) return bad;
        if (b.kind == BrickKind::Parted && b.data.size() != (size_t)t->brick_dim * t->brick_dim * t->brick_dim) return bad;
        t->node_mips.push_back(std::move(b));
    }
    t->node_mips.resize(t->nodes.len());
    // palettes and their lookup maps (bytecode.rs:640-655)
    if (f[6].kind != Item::List || f[7].kind != Item::List || f[8].kind != Item::List) return bad;
    for (size_t i = 0; i < f[6].list.size(); ++i) {
        const Item& a = f[6].list[i];
        if (a.kind != Item::List || a.list.size() != 4) return bad;
        for (const Item& c : a.list)
            if (c.kind != Item::Int || c.integer > 255) return bad;
        const Albedo al{(uint8_t)a.list[0].integer, (uint8_t)a.list[1].integer, (uint8_t)a.list[2].integer, (uint8_t)a.list[3].integer};
        t->voxel_color_palette.push_back(al);
        t->color_lookup_[((uint32_t)al.r << 24) | ((uint32_t)al.g << 16) | ((uint32_t)al.b << 8) | al.a] = i;
    }
    for (size_t i = 0; i < f[7].list.size(); ++i) {
        if (f[7].list[i].kind != Item::Int) return bad;
        t->voxel_data_palette.push_back((uint32_t)f[7].list[i].integer);
        t->data_lookup_[(uint32_t)f[7].list[i].integer] = i;
    }
    // MIPMapStrategy::decode_bencode_object (bytecode.rs:456-516), MIPResamplingMethods (:537-569)
    {
        const std::vector<Item>& m = f[8].list;
        size_t k = 0;
        auto next_int = [&](uint64_t* v) {
            if (k >= m.size() || m[k].kind != Item::Int) return false;
            *v = m[k++].integer;
            return true;
        };
        uint64_t enabled, n;
        if (!next_int(&enabled) || enabled > 1 || !next_int(&n)) return bad;
        t->mip_map_strategy.enabled = enabled == 1;
        t->mip_map_strategy.resampling_methods.clear();
        t->mip_map_strategy.resampling_color_matching_thresholds.clear();
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t level, code;
            if (!next_int(&level) || !next_int(&code)) return bad;
            MipSampler smp;
            if (code == 0) smp.method = MipMethod::BoxFilter;
            else if (code == 1) smp.method = MipMethod::PointFilter;
            else if (code == 2) smp.method = MipMethod::PointFilterBD;
            else if (code >= 3 && code < 1002) smp = MipSampler{MipMethod::Posterize, ((float)(uint32_t)code - 3.0f) / 1000.0f};
            else if (code >= 1003 && code < 2001) smp = MipSampler{MipMethod::PosterizeBD, ((float)(uint32_t)code - 1003.0f) / 1000.0f};
            else return bad;
            t->mip_map_strategy.resampling_methods[(size_t)level] = smp;
        }
        if (!next_int(&n)) return bad;
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t level, milli;
            if (!next_int(&level) || !next_int(&milli)) return bad;
            t->mip_map_strategy.resampling_color_matching_thresholds[(size_t)level] = (float)(uint32_t)milli / 1000.0f;
        }
        if (k != m.size()) return bad;
    }
    *out = hold.release();
    return OK;
}

}  // namespace svxo
