"""Host image of the uploaded node table (svx_octree_render_data_nodes): the flattening the reference does in
OctreeRenderData (src/raytracing/bevy/types.rs:216-279), in this library's 64-byte record layout (csrc/gpu_tree.hpp).
CPU only: the records are what the traversal kernels read, so their invariants are checked without a GPU -
the parent index and octant-in-parent a POP relies on (raytracing_on_cpu.rs:445-474), the bounds of
Cube::child_bounds_for (src/spatial/mod.rs:32-39), the occupied bits and brick kinds against Octree::get."""
import numpy as np
import pytest

import shocovox_b200 as S
from shocovox_b200 import scenes

NIL = 0xFFFFFFFF
NOTHING, INTERNAL, LEAF, UNIFORM = 0, 1, 2, 3


def records_of(tree):
    rec = tree.render_data_nodes()
    assert rec.dtype == np.uint32 and rec.shape[1] == 16
    return rec


def small_trees():
    yield "cpu_render", scenes.build_tree(scenes.cpu_render_scene(), S.Octree)  # the examples/cpu_render.rs scene
    deep = S.Octree(512, 1)  # deeper than the reference's 4-entry node stack (tests.rs:631)
    for i in range(0, 512, 37):
        deep.insert((i, (i * 7) % 512, (i * 13) % 512), S.Albedo(255, i % 256, 0, 255))
    yield "deep", deep
    mixed = S.Octree(128, 8)  # insert_at_lod slabs (Solid bricks, UniformLeaf nodes) next to per-voxel detail
    mixed.insert_at_lod((0, 0, 0), 32, S.Albedo(10, 20, 30, 255))
    mixed.insert_at_lod((64, 64, 64), 64, S.Albedo(40, 50, 60, 255))
    for i in range(40):
        mixed.insert((33 + i, 3, 5 + (i % 7)), S.Albedo(200, 100, i, 255))
    yield "mixed", mixed


@pytest.mark.parametrize("name,tree", list(small_trees()), ids=lambda v: v if isinstance(v, str) else "")
def test_records_are_a_consistent_breadth_first_tree(name, tree):
    rec = records_of(tree)
    n = len(rec)
    assert 1 <= n <= tree.node_count()  # only reachable nodes are serialised
    meta, parent = rec[:, 2], rec[:, 3]
    bounds = rec[:, 12:16].view(np.float32)
    size = float(tree.get_size())
    assert parent[0] == NIL and tuple(bounds[0]) == (0.0, 0.0, 0.0, size)
    seen_as_child = np.zeros(n, dtype=bool)
    order = []
    for i in range(n):
        kind = int(meta[i]) & 3
        if kind != INTERNAL:
            continue
        for o in range(8):
            c = int(rec[i, 4 + o])
            if c == NIL:
                continue
            assert 0 < c < n and not seen_as_child[c], "every node but the root has exactly one parent"
            seen_as_child[c] = True
            order.append(c)
            # what a POP reads: the parent's index and the octant the node occupies in it
            assert int(parent[c]) == i
            assert (int(meta[c]) >> 20) & 7 == o
            # Cube::child_bounds_for: octant bit 0 = x, bit 2 = y, bit 1 = z
            half = bounds[i, 3] / 2
            want = (bounds[i, 0] + (o & 1) * half, bounds[i, 1] + ((o >> 2) & 1) * half, bounds[i, 2] + ((o >> 1) & 1) * half, half)
            assert tuple(bounds[c]) == tuple(np.float32(v) for v in want)
    assert seen_as_child[1:].all() and not seen_as_child[0]
    assert order == sorted(order), "children are numbered breadth-first"


@pytest.mark.parametrize("name,tree", list(small_trees()), ids=lambda v: v if isinstance(v, str) else "")
def test_leaf_records_agree_with_get(name, tree):
    """Brick kinds and the occupied bits of leaf records against point queries: a brick that is not Empty covers at least
    one voxel that Octree::get reports, an Empty one none (sampled on a coarse lattice of each octant)."""
    rec = records_of(tree)
    bounds = rec[:, 12:16].view(np.float32)
    for i in range(len(rec)):
        meta = int(rec[i, 2])
        kind = meta & 3
        if kind not in (LEAF, UNIFORM):
            continue
        x0, y0, z0, s = (int(v) for v in bounds[i])
        octants = range(8) if kind == LEAF else [None]
        for o in octants:
            if o is None:
                bk, bx, by, bz, bs = (meta >> 2) & 3, x0, y0, z0, s
                slot = int(rec[i, 4])
            else:
                h = s // 2
                bk, bx, by, bz, bs = (meta >> (2 + 2 * o)) & 3, x0 + (o & 1) * h, y0 + ((o >> 2) & 1) * h, z0 + ((o >> 1) & 1) * h, h
                slot = int(rec[i, 4 + o])
            step = max(1, bs // 8)
            found = any(not tree.get((x, y, z)).is_none()
                        for x in range(bx, bx + bs, step) for y in range(by, by + bs, step) for z in range(bz, bz + bs, step))
            if bk == 0:
                assert slot == NIL and not found
            elif bk == 2:  # Solid: every voxel of the brick is the slot's palette value
                assert slot != NIL and found
        occupied = int(rec[i, 0]) | (int(rec[i, 1]) << 32)
        any_brick = any(((meta >> (2 + 2 * o)) & 3) != 0 for o in (range(8) if kind == LEAF else [0]))
        assert (occupied != 0) == any_brick


def test_count_query_and_small_buffer():
    import ctypes as C

    tree = scenes.build_tree(scenes.cpu_render_scene(), S.Octree)
    L = S.lib()
    n = C.c_uint64(0)
    assert L.svx_octree_render_data_nodes(tree._h, None, 0, C.byref(n)) == 0 and n.value > 1
    buf = (C.c_uint32 * 16)()
    assert L.svx_octree_render_data_nodes(tree._h, buf, 1, C.byref(n)) == 5  # SVX_E_INVALID_ARGUMENT: more than one node
    assert L.svx_octree_render_data_nodes(None, None, 0, C.byref(n)) == 5
