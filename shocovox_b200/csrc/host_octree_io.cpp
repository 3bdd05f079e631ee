// Bencode persistence of the host octree: Octree::to_bytes / from_bytes / save / load of the reference
// (src/octree/mod.rs:138-168) in the byte format its ToBencode / FromBencode impls define
// (src/convert/bytecode.rs, src/object_pool.rs:25-137), so a tree saved by the Rust crate can be rendered here and
// the other way round. `file:line` citations are relative to the reference checkout.
//
// Bencode (the `bendy` crate, Cargo.toml:19): integer = i<decimal>e, byte string = <len>:<bytes>, list = l<items>e.
//
//   Octree        = l i<auto_simplify> i<octree_size> i<brick_dim> NODES CHILDREN MIPS COLORS DATAS STRATEGY e   (bytecode.rs:579-598)
//   NODES         = l i<first_available> l ITEM* e e            ObjectPool (object_pool.rs:97-108)
//   ITEM          = l i<reserved> CONTENT e                     ReusableItem (object_pool.rs:25-36)
//   CONTENT       = 1:#  |  l 2:## i<occupied_bits> e  |  l 3:### BRICK x8 e  |  l 4:##u# BRICK e   (bytecode.rs:166-196)
//   BRICK         = 2:#b  |  l 3:#b# i<voxel> e  |  l 4:##b# i<len> i<voxel>*len 1:# e              (bytecode.rs:67-91)
//   CHILDREN      = l ( 5:##x##  |  l 5:##c## i<key> x8 e  |  l 5:##b## i<bitmap> e )* e            (bytecode.rs:289-311)
//   MIPS          = l BRICK* e                                  node_mips, one per node (types.rs:186)
//   COLORS        = l ( l i<r> i<g> i<b> i<a> e )* e            (bytecode.rs:11-22)
//   DATAS         = l i<u32>* e                                 Octree<T = u32>
//   STRATEGY      = l i<enabled> i<n> (i<level> i<method>)*n i<m> (i<level> i<threshold*1000>)*m e    (bytecode.rs:436-453)
//   method        = 0 BoxFilter | 1 PointFilter | 2 PointFilterBD | 3 + thr*1000 Posterize | 1003 + thr*1000 PosterizeBD
//                   (bytecode.rs:519-535; the decoder :537-569 accepts 3..1002 and 1003..2001 exclusive, so
//                   Posterize(0.999) cannot be read back and Posterize(1.0) comes back as PosterizeBD(0.0) - kept)
//
// The reference writes the two strategy maps in HashMap iteration order (random per process); we write them sorted
// by level.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "host_octree.hpp"

namespace svx {

namespace {

// Appends through a raw cursor into a malloc'd buffer (grown with realloc, handed to the caller as it is), head-room
// checked once per item. The voxel runs of Parted bricks are > 99 % of a file at ~12 bytes per voxel, and they repeat a
// handful of palette values: a small direct-mapped cache keeps the formatted text of the values seen last.
struct Writer {
    char* base = nullptr;
    char* cur = nullptr;
    char* lim = nullptr;
    bool failed = false;
    struct Slot {
        uint32_t value;
        uint32_t len;  // 0 = never filled
        char text[16];
    };
    Slot cache[1024];

    Writer() {
        for (Slot& c : cache) c.value = c.len = 0u;
    }
    ~Writer() { std::free(base); }
    // the finished buffer (never null on success; the caller frees it with std::free)
    uint8_t* release(size_t* len) {
        if (failed) return nullptr;
        need(1);
        if (failed) return nullptr;
        *len = (size_t)(cur - base);
        uint8_t* out = (uint8_t*)base;
        base = cur = lim = nullptr;
        return out;
    }
    bool need(size_t n) {
        if ((size_t)(lim - cur) >= n && base) return true;
        if (failed) return false;
        const size_t used = (size_t)(cur - base), cap = std::max(used + n, used + used / 2 + (size_t)65536);
        char* grown = (char*)std::realloc(base, cap);
        if (!grown) {
            failed = true;  // everything after this point is dropped; release() reports it
            return false;
        }
        base = grown;
        cur = grown + used;
        lim = grown + cap;
        return true;
    }
    static const char* pairs() {
        return "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
               "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
    }
    // "i<v>e" into out[0..16): ten zero-padded digits, then the 'i' goes over the last padding zero
    static uint32_t format_u32(uint32_t v, char* out) {
        const char* P = pairs();
        char buf[28];
        const uint32_t hi = v / 100000000u, rest = v - hi * 100000000u, mid = rest / 10000u, lo = rest - mid * 10000u;
        std::memcpy(buf + 1, P + 2 * hi, 2);
        std::memcpy(buf + 3, P + 2 * (mid / 100u), 2);
        std::memcpy(buf + 5, P + 2 * (mid % 100u), 2);
        std::memcpy(buf + 7, P + 2 * (lo / 100u), 2);
        std::memcpy(buf + 9, P + 2 * (lo % 100u), 2);
        buf[11] = 'e';
        const uint32_t nd = 1u + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) +
                            (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
        char* b = buf + 10 - nd;
        *b = 'i';
        std::memcpy(out, b, 16);  // nd + 2 <= 12 bytes are meaningful
        return nd + 2u;
    }
    void integer(uint64_t v) {
        if (!need(24)) return;
        if (v <= 0xFFFFFFFFull) {
            cur += format_u32((uint32_t)v, cur);
            return;
        }
        char buf[20];
        char* e = buf + 20;
        char* b = e;
        while (v >= 100) {
            const unsigned r = (unsigned)(v % 100);
            v /= 100;
            b -= 2;
            std::memcpy(b, pairs() + 2 * r, 2);
        }
        if (v >= 10) {
            b -= 2;
            std::memcpy(b, pairs() + 2 * v, 2);
        } else {
            *--b = (char)('0' + v);
        }
        *cur++ = 'i';
        std::memcpy(cur, b, (size_t)(e - b));
        cur += e - b;
        *cur++ = 'e';
    }
    void integers(const uint32_t* v, size_t n) {  // i<v0>e i<v1>e ...: at most 12 bytes each
        if (!need(n * 12 + 16)) return;
        for (size_t k = 0; k < n; ++k) {
            const uint32_t x = v[k];
            Slot& c = cache[(x ^ ((x >> 16) * 0x9E3779B1u)) & 1023u];
            if (c.value != x || c.len == 0u) {
                c.value = x;
                c.len = format_u32(x, c.text);
            }
            std::memcpy(cur, c.text, 16);
            cur += c.len;
        }
    }
    void str(const char* lit) {
        const size_t n = std::strlen(lit);
        if (!need(n + 24)) return;
        const uint32_t d = format_u32((uint32_t)n, cur);  // "i<n>e" -> "<n>:"
        std::memmove(cur, cur + 1, d - 2);
        cur += d - 2;
        *cur++ = ':';
        std::memcpy(cur, lit, n);
        cur += n;
    }
    void open() {
        if (need(1)) *cur++ = 'l';
    }
    void close() {
        if (need(1)) *cur++ = 'e';
    }
};

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;

    bool fail() {
        ok = false;
        return false;
    }
    bool peek_list() const { return p < end && *p == 'l'; }
    bool peek_int() const { return p < end && *p == 'i'; }
    bool peek_bytes() const { return p < end && *p >= '0' && *p <= '9'; }
    bool at_close() const { return p < end && *p == 'e'; }
    bool open() {
        if (!peek_list()) return fail();
        ++p;
        return true;
    }
    bool close() {
        if (!at_close()) return fail();
        ++p;
        return true;
    }
    bool integer(uint64_t* out) {
        if (!peek_int()) return fail();
        const uint8_t* first = p + 1;
        const uint8_t* q = first;
        uint64_t v = 0;
        while (q < end && (unsigned)(*q - '0') <= 9u) {  // up to 19 digits cannot overflow; longer ones are re-read below
            v = v * 10 + (uint64_t)(*q - '0');
            ++q;
        }
        const size_t digits = (size_t)(q - first);
        if (digits == 0 || q >= end || *q != 'e') return fail();  // no negative numbers anywhere in the format
        if (digits > 19) {
            v = 0;
            for (const uint8_t* c = first; c < q; ++c) {
                const uint64_t d = (uint64_t)(*c - '0');
                if (v > (UINT64_MAX - d) / 10) return fail();
                v = v * 10 + d;
            }
        }
        p = q + 1;
        *out = v;
        return true;
    }
    bool bytes(const char** s, size_t* n) {
        if (!peek_bytes()) return fail();
        size_t len = 0;
        while (p < end && *p >= '0' && *p <= '9') {
            if (len > (SIZE_MAX - 9) / 10) return fail();
            len = len * 10 + (size_t)(*p - '0');
            ++p;
        }
        if (p >= end || *p != ':') return fail();
        ++p;
        if ((size_t)(end - p) < len) return fail();
        *s = (const char*)p;
        *n = len;
        p += len;
        return true;
    }
    bool marker(const char* lit) {
        const char* s;
        size_t n;
        if (!bytes(&s, &n)) return false;
        if (n != std::strlen(lit) || std::memcmp(s, lit, n) != 0) return fail();
        return true;
    }
    // skips one item of any type
    bool skip() {
        if (peek_int()) {
            uint64_t v;
            return integer(&v);
        }
        if (peek_bytes()) {
            const char* s;
            size_t n;
            return bytes(&s, &n);
        }
        if (!open()) return false;
        while (ok && !at_close()) {
            if (p >= end) return fail();
            if (!skip()) return false;
        }
        return close();
    }
};

bool is_marker(const char* s, size_t n, const char* lit) { return n == std::strlen(lit) && std::memcmp(s, lit, n) == 0; }

}  // namespace

// ---- encode ---------------------------------------------------------------------------------------------------
uint8_t* HostOctree::to_bytes(size_t* len) const {
    std::unique_ptr<Writer> writer(new (std::nothrow) Writer());  // 24 KB of cache: not on the stack
    if (!writer) return nullptr;
    Writer& w = *writer;
    auto brick = [&](const BrickRef& b) {  // bytecode.rs:67-91
        if (b.kind == BK_EMPTY) {
            w.str("#b");
        } else if (b.kind == BK_SOLID) {
            w.open();
            w.str("#b#");
            w.integer(b.value);
            w.close();
        } else {
            const uint32_t* v = brick_data(b.value);
            w.open();
            w.str("##b#");
            w.integer(vol_);
            w.integers(v, vol_);
            w.str("#");
            w.close();
        }
    };
    w.open();
    w.integer(auto_simplify ? 1 : 0);
    w.integer(size_);
    w.integer(dim_);
    // nodes: ObjectPool { first_available, buffer } (object_pool.rs:97-108)
    w.open();
    w.integer(first_available_);
    w.open();
    for (const NodeRec& n : nodes_) {
        w.open();
        w.integer(n.reserved ? 1 : 0);
        switch (n.kind) {  // bytecode.rs:166-196
            case NK_INTERNAL:
                w.open();
                w.str("##");
                w.integer(n.ocbits);
                w.close();
                break;
            case NK_LEAF:
                w.open();
                w.str("###");
                for (int o = 0; o < 8; ++o) brick(n.brick[o]);
                w.close();
                break;
            case NK_UNIFORM:
                w.open();
                w.str("##u#");
                brick(n.brick[0]);
                w.close();
                break;
            default: w.str("#"); break;
        }
        w.close();
    }
    w.close();
    w.close();
    // node_children (bytecode.rs:289-311)
    w.open();
    for (const NodeRec& n : nodes_) {
        if (n.link == LK_CHILDREN) {
            w.open();
            w.str("##c##");
            for (int o = 0; o < 8; ++o) w.integer(n.child[o]);
            w.close();
        } else if (n.link == LK_BITMAP) {
            w.open();
            w.str("##b##");
            w.integer(n.leaf_bits);
            w.close();
        } else {
            w.str("##x##");
        }
    }
    w.close();
    // node_mips: one brick per node key (bytecode.rs:591)
    w.open();
    for (const NodeRec& n : nodes_) brick(n.mip);
    w.close();
    w.open();
    for (const svx_albedo& a : colors_) {  // bytecode.rs:11-22
        w.open();
        w.integer(a.r);
        w.integer(a.g);
        w.integer(a.b);
        w.integer(a.a);
        w.close();
    }
    w.close();
    w.open();
    for (uint32_t d : datas_) w.integer(d);
    w.close();
    // MIPMapStrategy (bytecode.rs:436-453): thresholds and Posterize parameters as `(thr * 1000.) as u32`
    auto milli = [](float thr) -> uint64_t {
        const float v = thr * 1000.0f;
        if (!(v > 0.0f)) return 0;
        return v >= 4294967296.0f ? 0xFFFFFFFFull : (uint64_t)(uint32_t)v;
    };
    w.open();
    w.integer(mips_enabled_ ? 1 : 0);
    w.integer(mip_methods_.size());
    for (const auto& m : mip_methods_) {
        w.integer(m.first);
        switch (m.second.method) {  // bytecode.rs:519-535
            case MIP_POSTERIZE: w.integer(3 + milli(m.second.thr)); break;
            case MIP_POSTERIZE_BD: w.integer(1003 + milli(m.second.thr)); break;
            default: w.integer(m.second.method); break;
        }
    }
    w.integer(mip_thresholds_.size());
    for (const auto& t : mip_thresholds_) {
        w.integer(t.first);
        w.integer(milli(t.second));
    }
    w.close();
    w.close();
    return w.release(len);
}

// ---- decode ---------------------------------------------------------------------------------------------------
int32_t HostOctree::from_bytes(const uint8_t* data, size_t len, HostOctree** out) {
    *out = nullptr;
    if (!data) return SVX_E_DECODE;
    Reader r{data, data + len};
    uint64_t auto_simplify = 0, size = 0, dim = 0;
    if (!r.open() || !r.integer(&auto_simplify) || !r.integer(&size) || !r.integer(&dim)) return SVX_E_DECODE;
    if (auto_simplify > 1 || size > 0xFFFFFFFFull || dim > 0xFFFFFFFFull) return SVX_E_DECODE;
    HostOctree* t = nullptr;
    const int32_t created = create((uint32_t)size, (uint32_t)dim, &t);  // the same validation as Octree::new
    if (created != SVX_OK) return created;
    struct Guard {
        HostOctree* t;
        ~Guard() { delete t; }
    } guard{t};
    t->auto_simplify = auto_simplify == 1;
    t->nodes_.clear();

    auto brick = [&](BrickRef* b) -> bool {  // bytecode.rs:94-160
        b->kind = BK_EMPTY;
        b->value = NIL;
        if (r.peek_bytes()) return r.marker("#b");
        if (!r.open()) return false;
        const char* s;
        size_t n;
        if (!r.bytes(&s, &n)) return false;
        if (is_marker(s, n, "#b#")) {
            uint64_t v;
            if (!r.integer(&v) || v > 0xFFFFFFFFull) return r.fail();
            b->kind = BK_SOLID;
            b->value = (uint32_t)v;
        } else if (is_marker(s, n, "##b#")) {
            uint64_t count;
            if (!r.integer(&count) || count != t->vol_) return r.fail();
            const uint32_t h = t->brick_alloc(NIL);
            b->kind = BK_PARTED;
            b->value = h;
            uint32_t* v = t->brick_mut(h);
            for (uint32_t i = 0; i < t->vol_; ++i) {
                uint64_t x;
                if (!r.integer(&x) || x > 0xFFFFFFFFull) return r.fail();
                v[i] = (uint32_t)x;
            }
            if (r.peek_bytes() && !r.marker("#")) return false;  // trailing "#" (the decoder of the reference ignores it)
        } else {
            return r.fail();
        }
        return r.close();
    };

    // nodes
    uint64_t first_available = 0;
    if (!r.open() || !r.integer(&first_available) || !r.open()) return SVX_E_DECODE;
    while (r.ok && !r.at_close()) {
        NodeRec n;
        uint64_t reserved;
        if (!r.open() || !r.integer(&reserved) || reserved > 1) return SVX_E_DECODE;
        n.reserved = (uint8_t)reserved;
        if (r.peek_bytes()) {
            if (!r.marker("#")) return SVX_E_DECODE;
            n.kind = NK_NOTHING;
        } else {
            const char* s;
            size_t k;
            if (!r.open() || !r.bytes(&s, &k)) return SVX_E_DECODE;
            if (is_marker(s, k, "##")) {
                n.kind = NK_INTERNAL;
                if (!r.integer(&n.ocbits)) return SVX_E_DECODE;
            } else if (is_marker(s, k, "###")) {
                n.kind = NK_LEAF;
                for (int o = 0; o < 8; ++o)
                    if (!brick(&n.brick[o])) return SVX_E_DECODE;
            } else if (is_marker(s, k, "##u#")) {
                n.kind = NK_UNIFORM;
                if (!brick(&n.brick[0])) return SVX_E_DECODE;
            } else {
                return SVX_E_DECODE;
            }
            if (!r.close()) return SVX_E_DECODE;
        }
        if (!r.close()) return SVX_E_DECODE;
        if (!n.reserved) {
            // ObjectPool::free leaves the dead item in an un-reserved slot (object_pool.rs:213-221) and the reference writes
            // it out; nothing ever reads it (push overwrites it, :172-176). This host tree keeps freed slots empty.
            for (auto& b : n.brick) t->brick_release(b);
            n.kind = NK_NOTHING;
            n.ocbits = 0;
        }
        t->nodes_.push_back(n);
    }
    if (!r.close() || !r.close() || t->nodes_.empty()) return SVX_E_DECODE;
    t->first_available_ = (size_t)std::min<uint64_t>(first_available, t->nodes_.size());
    // node_children, parallel to the node buffer
    if (!r.open()) return SVX_E_DECODE;
    size_t ci = 0;
    while (r.ok && !r.at_close()) {
        NodeRec scratch;
        NodeRec& n = ci < t->nodes_.size() ? t->nodes_[ci] : scratch;
        ++ci;
        if (r.peek_bytes()) {
            if (!r.marker("##x##")) return SVX_E_DECODE;
            n.link = LK_NONE;
            continue;
        }
        const char* s;
        size_t k;
        if (!r.open() || !r.bytes(&s, &k)) return SVX_E_DECODE;
        if (is_marker(s, k, "##c##")) {
            n.link = LK_CHILDREN;
            for (int o = 0; o < 8; ++o) {
                uint64_t c;
                if (!r.integer(&c) || c > 0xFFFFFFFFull) return SVX_E_DECODE;
                n.child[o] = (uint32_t)c;
            }
        } else if (is_marker(s, k, "##b##")) {
            n.link = LK_BITMAP;
            if (!r.integer(&n.leaf_bits)) return SVX_E_DECODE;
        } else {
            return SVX_E_DECODE;
        }
        if (!r.close()) return SVX_E_DECODE;
    }
    if (!r.close()) return SVX_E_DECODE;
    // node_mips, parallel to the node buffer (bytecode.rs:638)
    if (!r.open()) return SVX_E_DECODE;
    size_t mi = 0;
    while (r.ok && !r.at_close()) {
        BrickRef scratch;
        BrickRef* m = mi < t->nodes_.size() ? &t->nodes_[mi].mip : &scratch;
        ++mi;
        if (!brick(m)) return SVX_E_DECODE;
        if (m == &scratch) t->brick_release(scratch);
    }
    if (!r.close()) return SVX_E_DECODE;
    // palettes (bytecode.rs:640-655: later duplicates win in the lookup maps)
    if (!r.open()) return SVX_E_DECODE;
    while (r.ok && !r.at_close()) {
        uint64_t c[4];
        if (!r.open()) return SVX_E_DECODE;
        for (auto& v : c)
            if (!r.integer(&v) || v > 255) return SVX_E_DECODE;
        if (!r.close()) return SVX_E_DECODE;
        const svx_albedo a{(uint8_t)c[0], (uint8_t)c[1], (uint8_t)c[2], (uint8_t)c[3]};
        t->color_index_[((uint32_t)a.r << 24) | ((uint32_t)a.g << 16) | ((uint32_t)a.b << 8) | a.a] = (uint32_t)t->colors_.size();
        t->colors_.push_back(a);
    }
    if (!r.close() || !r.open()) return SVX_E_DECODE;
    while (r.ok && !r.at_close()) {
        uint64_t d;
        if (!r.integer(&d) || d > 0xFFFFFFFFull) return SVX_E_DECODE;
        t->data_index_[(uint32_t)d] = (uint32_t)t->datas_.size();
        t->datas_.push_back((uint32_t)d);
    }
    if (!r.close()) return SVX_E_DECODE;
    // MIPMapStrategy (bytecode.rs:456-516)
    {
        uint64_t enabled = 0, count = 0;
        if (!r.open() || !r.integer(&enabled) || enabled > 1 || !r.integer(&count)) return SVX_E_DECODE;
        std::map<size_t, MipSampler> methods;
        for (uint64_t i = 0; i < count; ++i) {
            uint64_t level, code;
            if (!r.integer(&level) || !r.integer(&code) || code > 0xFFFFFFFFull) return SVX_E_DECODE;
            MipSampler m;
            if (code <= 2) {
                m.method = (uint32_t)code;
            } else if (code >= 3 && code < 1002) {
                m.method = MIP_POSTERIZE;
                m.thr = ((float)(uint32_t)code - 3.0f) / 1000.0f;
            } else if (code >= 1003 && code < 2001) {
                m.method = MIP_POSTERIZE_BD;
                m.thr = ((float)(uint32_t)code - 1003.0f) / 1000.0f;
            } else {
                return SVX_E_DECODE;
            }
            methods[(size_t)level] = m;
        }
        if (!r.integer(&count)) return SVX_E_DECODE;
        std::map<size_t, float> thresholds;
        for (uint64_t i = 0; i < count; ++i) {
            uint64_t level, milli;
            if (!r.integer(&level) || !r.integer(&milli) || milli > 0xFFFFFFFFull) return SVX_E_DECODE;
            thresholds[(size_t)level] = (float)(uint32_t)milli / 1000.0f;
        }
        if (!r.close()) return SVX_E_DECODE;
        t->mip_load_strategy(enabled == 1, std::move(methods), std::move(thresholds));
    }
    if (!r.close() || r.p != r.end) return SVX_E_DECODE;
    if (t->colors_.size() > 0xFFFF || t->datas_.size() > 0xFFFF) return SVX_E_DECODE;  // u16 palette indices, types.rs:188-191
    // Untrusted input: the reference would bounds-panic on a palette index beyond its palette and recurse forever on a
    // child cycle. Reject both here so that get(), the serialiser and the kernels can trust the tree.
    auto value_ok = [&](uint32_t v) {
        const uint32_t ci = v & 0xFFFFu, di = v >> 16;
        return (ci == 0xFFFFu || ci < t->colors_.size()) && (di == 0xFFFFu || di < t->datas_.size());
    };
    auto brick_ok = [&](const BrickRef& b) {
        if (b.kind == BK_SOLID) return value_ok(b.value);
        if (b.kind == BK_PARTED) {
            const uint32_t* v = t->brick_data(b.value);
            for (uint32_t i = 0; i < t->vol_; ++i)
                if (!value_ok(v[i])) return false;
        }
        return true;
    };
    for (const NodeRec& n : t->nodes_) {
        if (!brick_ok(n.mip)) return SVX_E_DECODE;  // a MIP belongs to the key and survives a freed slot
        if (!n.reserved) continue;
        // content and connection must be of one kind: leaves carry an occupancy bitmap, internal nodes child keys (the
        // reference only debug-asserts this, e.g. detail.rs:524-544; a file that disagrees would be walked as the wrong thing)
        if (((n.kind == NK_LEAF || n.kind == NK_UNIFORM) && n.link == LK_CHILDREN) || (n.kind == NK_INTERNAL && n.link == LK_BITMAP))
            return SVX_E_DECODE;
        for (const BrickRef& b : n.brick)
            if (!brick_ok(b)) return SVX_E_DECODE;
    }
    {
        std::vector<uint8_t> seen(t->nodes_.size(), 0);
        std::vector<std::pair<uint32_t, uint32_t>> todo{{0u, t->size_}};  // (key, node size)
        seen[0] = 1;
        while (!todo.empty()) {
            const auto [key, node_size] = todo.back();
            todo.pop_back();
            const NodeRec& n = t->nodes_[key];
            if (n.kind != NK_INTERNAL || n.link != LK_CHILDREN) continue;
            for (uint32_t c : n.child) {
                if (!t->key_is_valid(c)) continue;
                if (seen[c] || node_size / 2 < t->dim_) return SVX_E_DECODE;  // shared / cyclic child, or below brick size (a UniformLeaf can be as small as one brick)
                seen[c] = 1;
                todo.push_back({c, node_size / 2});
            }
        }
    }
    t->revision_ += 1;
    guard.t = nullptr;
    *out = t;
    return SVX_OK;
}

int32_t HostOctree::save(const char* path) const {  // octree/mod.rs:144-150
    size_t len = 0;
    uint8_t* bytes = to_bytes(&len);
    if (!bytes) return SVX_E_OUT_OF_MEMORY;
    std::FILE* f = std::fopen(path, "wb");
    if (!f) {
        std::free(bytes);
        return SVX_E_IO;
    }
    const bool ok = std::fwrite(bytes, 1, len, f) == len;
    std::free(bytes);
    return (std::fclose(f) == 0 && ok) ? SVX_OK : SVX_E_IO;
}

int32_t HostOctree::load(const char* path, HostOctree** out) {  // octree/mod.rs:153-159
    *out = nullptr;
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return SVX_E_IO;
    std::string bytes;
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) bytes.append(buf, n);
    const bool ok = !std::ferror(f);
    std::fclose(f);
    if (!ok) return SVX_E_IO;
    return from_bytes((const uint8_t*)bytes.data(), bytes.size(), out);
}

}  // namespace svx
