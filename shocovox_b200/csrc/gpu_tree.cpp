// Breadth-first serialisation of the host octree into the device layout (see gpu_tree.hpp).
#include "gpu_tree.hpp"

#include <cstring>

namespace svx {

void serialise_nodes(const HostOctree& tree, SerialisedNodes* out) {
    SerialisedNodes& s = *out;
    s = SerialisedNodes();
    s.tree_size = tree.size();
    s.brick_dim = tree.brick_dim();
    s.brick_shift = 0;
    while ((1u << s.brick_shift) < s.brick_dim) ++s.brick_shift;
    const uint32_t vol = tree.brick_volume();
    s.bit_words = (vol + 31) / 32;
    s.revision = tree.revision();

    const std::vector<NodeRec>& nodes = tree.nodes();
    // breadth-first order over reachable nodes; `order[i]` = host key of device node i
    std::vector<uint32_t> order;
    std::vector<uint32_t> level_end;
    order.push_back(0);
    for (size_t i = 0; i < order.size(); ++i) {
        const NodeRec& n = nodes[order[i]];
        if (n.kind != NK_INTERNAL || n.link != LK_CHILDREN) continue;
        for (int o = 0; o < 8; ++o)
            if (tree.key_is_valid(n.child[o])) order.push_back(n.child[o]);
    }
    std::vector<uint32_t> remap(nodes.size(), NIL);
    for (size_t i = 0; i < order.size(); ++i) remap[order[i]] = (uint32_t)i;

    s.node_head.resize(order.size());
    s.node_slot.assign(order.size() * 8, NIL);
    s.node_mip.assign(order.size(), NIL);
    s.mips_enabled = tree.mips_enabled();
    auto slot_of = [&](const BrickRef& b) -> uint32_t {
        if (b.kind == BK_PARTED) ++s.live_bricks;
        return b.kind == BK_EMPTY ? NIL : b.value;  // Solid: the palette value ; Parted: the pool handle
    };

    for (size_t i = 0; i < order.size(); ++i) {
        const uint32_t key = order[i];
        const NodeRec& n = nodes[key];
        const uint64_t oc = tree.stored_occupied_bits(key);
        NodeHead h{(uint32_t)oc, (uint32_t)(oc >> 32), n.kind, NIL};
        uint32_t* slot = s.node_slot.data() + i * 8;
        if (n.kind == NK_INTERNAL) {
            if (n.link == LK_CHILDREN)
                for (int o = 0; o < 8; ++o) slot[o] = tree.key_is_valid(n.child[o]) ? remap[n.child[o]] : NIL;
        } else if (n.kind == NK_LEAF) {
            for (int o = 0; o < 8; ++o) {
                h.meta |= (uint32_t)n.brick[o].kind << (2 + 2 * o);
                slot[o] = slot_of(n.brick[o]);
            }
        } else if (n.kind == NK_UNIFORM) {
            h.meta |= (uint32_t)n.brick[0].kind << 2;
            slot[0] = slot_of(n.brick[0]);
        }
        if (s.mips_enabled) {
            h.meta |= (uint32_t)n.mip.kind << 18;
            s.node_mip[i] = slot_of(n.mip);
        }
        s.node_head[i] = h;
    }
    // bounds and parent of every node, parents before children (breadth-first order): Cube::child_bounds_for,
    // src/spatial/mod.rs:32-39. Every node but the root has exactly one parent (from_bytes rejects shared children).
    s.node_bounds.assign(order.size() * 4, 0.0f);
    s.node_bounds[3] = (float)s.tree_size;
    for (size_t i = 0; i < order.size(); ++i) {
        if ((s.node_head[i].meta & 3u) != NK_INTERNAL) continue;
        const float half = s.node_bounds[i * 4 + 3] * 0.5f;
        for (int o = 0; o < 8; ++o) {
            const uint32_t c = s.node_slot[i * 8 + o];
            if (c == NIL) continue;
            s.node_head[c].aux = (uint32_t)i;
            s.node_head[c].meta |= (uint32_t)o << 20;  // the octant this node occupies in its parent
            s.node_bounds[c * 4 + 0] = s.node_bounds[i * 4 + 0] + (float)(o & 1) * half;         // octant bit 0: x
            s.node_bounds[c * 4 + 1] = s.node_bounds[i * 4 + 1] + (float)((o >> 2) & 1) * half;  // bit 2: y
            s.node_bounds[c * 4 + 2] = s.node_bounds[i * 4 + 2] + (float)((o >> 1) & 1) * half;  // bit 1: z
            s.node_bounds[c * 4 + 3] = half;
        }
    }
    // depth of the deepest node (root = 1): bounds the ring-stack overflow behaviour, reported in stats
    {
        std::vector<uint32_t> d(order.size(), 1);
        uint32_t deepest = 1;
        for (size_t i = 0; i < order.size(); ++i) {
            if ((s.node_head[i].meta & 3u) != NK_INTERNAL) continue;
            for (int o = 0; o < 8; ++o) {
                const uint32_t c = s.node_slot[i * 8 + o];
                if (c != NIL) {
                    d[c] = d[i] + 1;
                    deepest = d[c] > deepest ? d[c] : deepest;
                }
            }
        }
        s.depth = deepest;
    }
    const std::vector<svx_albedo>& pal = tree.color_palette();
    s.palette.resize(pal.size() ? pal.size() : 1, 0u);
    for (size_t i = 0; i < pal.size(); ++i)
        s.palette[i] = (uint32_t)pal[i].r | ((uint32_t)pal[i].g << 8) | ((uint32_t)pal[i].b << 16) | ((uint32_t)pal[i].a << 24);
}

}  // namespace svx
