// Device layout of the octree ("render data") and its breadth-first serialiser.
//
// The reference uploads 4-byte granular arrays on demand (metadata / node_children / node_ocbits / voxels,
// reference src/raytracing/bevy/types.rs:216-279, bevy/cache.rs:413-539). Here the whole tree is resident and a
// node visit is ONE 16-byte load, the three per-node tables below are interleaved into one 64-byte record per node
// (DeviceTree::node_rec):
//
//   node_head[i] : uint4 { ocbits.lo, ocbits.hi, meta, aux }
//        ocbits = stored_occupied_bits(node)   (reference src/octree/detail.rs:524-544, resolved on the host)
//        meta   = kind[1:0] | brick kind of octant o at [2+2o+1 : 2+2o]   (0 empty, 1 parted, 2 solid)
//                 kind: 0 Nothing, 1 Internal, 2 Leaf, 3 UniformLeaf (its brick kind sits in octant 0's field)
//                 bits [19:18]: kind of the node's MIP brick (reference node_mips[key], src/octree/types.rs:186)
//                 bits [22:20]: the octant this node occupies in its parent (what a POP derives from the two bounds,
//                               raytracing_on_cpu.rs:458-463)
//        aux    = index of the parent node (NIL for the root). The reference keeps the path in a 4-entry ring stack
//                 (NodeStack, raytracing_on_cpu.rs:20-82); its entries are always the current node's nearest ancestors, so
//                 a POP needs the parent index and a count of valid entries, not the entries themselves
//   node_slot[8i + o] : u32   Internal: child node index or NIL (validity resolved on the host)
//                             Leaf: brick slot of octant o
//                             UniformLeaf: slot 0 = the brick slot of its one brick
//        brick slot = palette value (Solid) | brick handle (Parted) | NIL (Empty)
//   node_mip[i] : u32   brick slot of the node's MIP brick; read only by the level-of-detail branch of
//                       get_by_ray_at_lod (raytracing_on_cpu.rs:368-386), and only present in the layout's hot loop
//                       when the tree's MIP maps are enabled (mips_enabled selects the kernel variant)
//   voxels[brick * dim^3 + x + y*dim + z*dim^2] : u32 palette values (reference flat_projection order)
//   brick_bits[brick * words + ...] : 1 bit per voxel, set when the voxel is NOT empty (pix_points_to_empty false);
//        the DDA walks these bits and fetches the 4-byte voxel only for the hit. Computed ON THE DEVICE from the
//        uploaded voxels (kernels.cu: occupancy_bits_kernel)
//   palette[c] : RGBA8, r in the low byte
//
// Nodes are numbered breadth-first from the root (index 0), so the top levels share cache lines; they are a few
// hundred KB at most and are re-serialised on every reload. Bricks are the bulk (C3: 550 MB) and are NOT renumbered:
// the device brick index is the host pool handle, so `voxels` mirrors the host's pooled voxel array one to one and a
// reload after an edit uploads only the bricks written since the last upload (HostOctree::brick_revision) - the job
// the reference's streaming cache does per node request (src/raytracing/bevy/data.rs:365-773).
#pragma once
#include <cstdint>
#include <vector>

#include <vector_types.h>

#include "host_octree.hpp"

namespace svx {

struct NodeHead {
    uint32_t oc_lo, oc_hi, meta, aux;
};

// POD handed to kernels by value
struct DeviceTree {
    // One 64-byte record per node, so that everything a node visit touches shares a cache line and one address:
    //   uint4[0] = node_head   uint4[1..2] = node_slot[0..7]   uint4[3] = node_bounds {min x, min y, min z, size} as f32
    // (the bounds are what a POP restores; exact integers in f32)
    const uint4* node_rec;
    const uint32_t* node_mip;
    const uint32_t* voxels;
    const uint32_t* brick_bits;
    const uint32_t* palette;
    // RAY_TO_NODE_OCCUPANCY_BITMASK_LUT (reference src/spatial/lut.rs, generator :39-89) as [direction octant][cell] uint2
    // {lo, hi}: 4 KB, identical for every tree, resident in L1. One 8-byte load replaces ~30 integer instructions of the
    // closed form (traverse.cuh: ray_may_hit), which stays as the start-up cross-check of this very table.
    const uint2* ray_lut;
    uint32_t n_nodes;
    uint32_t n_bricks;
    uint32_t tree_size;
    uint32_t brick_dim;
    uint32_t brick_shift;       // log2(brick_dim)
    uint32_t brick_dim_sq;      // brick_dim^2: the z stride of flat_projection
    uint32_t bit_words;         // u32 words of brick_bits per brick
    uint32_t n_colors;
    uint32_t mips_enabled;      // MIPMapStrategy::is_enabled (mipmap.rs:693): the LOD kernel variant runs
    float inv_tree_size;        // 1 / tree_size (exact, power of two)
    float inv_brick_dim;        // 1 / brick_dim (exact, power of two)
};

// Host image of the node part of the device layout
struct SerialisedNodes {
    std::vector<NodeHead> node_head;
    std::vector<uint32_t> node_slot;
    std::vector<uint32_t> node_mip;
    std::vector<float> node_bounds;  // 4 per node: min x, y, z, size
    std::vector<uint32_t> palette;
    uint32_t tree_size = 0, brick_dim = 0, brick_shift = 0, bit_words = 0, depth = 0;
    uint64_t live_bricks = 0;  // parted bricks referenced by reachable nodes (MIP bricks included)
    bool mips_enabled = false;
    uint64_t revision = 0;
};

void serialise_nodes(const HostOctree& tree, SerialisedNodes* out);

}  // namespace svx
