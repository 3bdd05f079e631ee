"""Bencode persistence (Octree::to_bytes / from_bytes / save / load, src/octree/mod.rs:138-168; byte format of
src/convert/bytecode.rs and src/object_pool.rs:25-137).

The reference's own tests (src/convert/bytecode_tests.rs) are round trips without literal bytes, and the Rust crate
cannot run here, so the format is pinned three ways: (1) a literal byte string assembled BY HAND from the reference's
encode functions for a one-voxel tree, (2) a third, minimal encoder written in this file from the same grammar, and
(3) two independent codecs - the oracle's (document-tree based, oracle/svx_oracle_bytecode.cpp) and the product's
(streaming, csrc/host_octree_io.cpp) - that must produce identical bytes and read each other's output.
Paths relative to /root/reference/.
"""
import numpy as np
import pytest

import oracle_lib as O
import shocovox_b200 as S
from oracle_lib import OracleOctree, entry_key as K
from product_adapter import ProductOctree
from shocovox_b200 import scenes

NIL = 0xFFFFFFFF


class OracleIO:
    to_bytes = staticmethod(lambda t: t.to_bytes())
    from_bytes = staticmethod(OracleOctree.from_bytes)
    new = OracleOctree


class ProductIO:
    to_bytes = staticmethod(lambda t: t.tree.to_bytes())
    new = ProductOctree

    @staticmethod
    def from_bytes(data):
        p = ProductOctree.__new__(ProductOctree)
        try:
            p.tree = S.Octree.from_bytes(data)
        except S.OctreeError as e:
            raise ValueError(e.code)
        return p


@pytest.fixture(params=["oracle", "product"])
def IO(request):
    return OracleIO if request.param == "oracle" else ProductIO


# ---- (1) hand-assembled literal ---------------------------------------------------------------------------------
# Octree::new(4, 1) + insert((1,2,3), Albedo(10,20,30,255)); traced through the reference by hand:
#  * root (key 0) becomes Internal; (1,2,3) is in root octant x=0,z=1,y=1 -> 0+2+4 = 6, so child[6] = key 1
#    (insert.rs:130-204); the root's occupied bits hold bitmap cell (1,2,3) of the 4x4x4 map: bit 1 + 4*2 + 16*3 = 57
#  * node 1 covers (0..2, 2..4, 2..4) and is a Leaf at brick_dim 1: the voxel sits in its octant x=1,z=1,y=0 -> 3,
#    BrickData::Solid(pix_visual(0)) = colour index 0, no data index = 0xFFFF0000 (node.rs:354-403)
#  * its 64-bit leaf bitmap marks the 2x2x2 cells of octant (1,0,1): x in 2..4, y in 0..2, z in 2..4
#    -> bits {2,3,6,7} + 32 and + 48 = 0x00CC00CC00000000 = 57421771422302208 (set_occupancy_in_bitmap_64bits)
#  * ObjectPool.first_available ends at 1 (object_pool.rs:178-202)
ONE_VOXEL = (
    b"l"
    b"i1e" b"i4e" b"i1e"                                    # auto_simplify, octree_size, brick_dim
    b"l" b"i1e"                                             # ObjectPool: first_available
    b"l"
    b"l" b"i1e" b"l2:##i144115188075855872ee" b"e"          # node 0: reserved, Internal(1 << 57)
    b"l" b"i1e" b"l3:###" b"2:#b2:#b2:#b" b"l3:#b#i4294901760ee" b"2:#b2:#b2:#b2:#b" b"e" b"e"  # node 1: Leaf
    b"e" b"e"
    b"l"
    b"l5:##c##" + b"i4294967295e" * 6 + b"i1e" + b"i4294967295e" + b"e"   # Children([MAX x6, 1, MAX])
    b"l5:##b##i57421771422302208ee"                                         # OccupancyBitmap
    b"e"
    b"l2:#b2:#be"                                           # node_mips
    b"ll" b"i10ei20ei30ei255e" b"ee"                        # colour palette
    b"le"                                                   # data palette
    b"l" b"i0e" b"i4e" b"i1ei1e" b"i2ei0e" b"i3ei0e" b"i4ei0e" b"i3e" b"i2ei100e" b"i3ei50e" b"i4ei20e" b"e"  # MIPMapStrategy::default()
    b"e"
)


def test_one_voxel_tree_matches_the_hand_assembled_bytes(IO):
    t = IO.new(4, 1)
    assert t.insert((1, 2, 3), (10, 20, 30, 255)) == O.OK
    assert IO.to_bytes(t) == ONE_VOXEL
    u = IO.from_bytes(ONE_VOXEL)
    assert u.get((1, 2, 3)) == K((10, 20, 30, 255))
    assert sum(u.get((x, y, z)) != K() for x in range(4) for y in range(4) for z in range(4)) == 1
    assert u.structure_hash() == t.structure_hash()


# ---- (1b) a second, larger hand-assembled literal: two levels, Parted + Solid bricks, both palettes ----------------------------
# Octree::new(8, 2) then
#   insert((0,0,0), Visual(255,0,0,255))            -> colour 0
#   insert((1,1,1), Complex((0,255,0,255), 7))      -> colour 1, data 0
#   insert_at_lod((4,4,4), 4, Visual(0,255,0,255))  -> colour 1 again (add_to_palette finds it, update/mod.rs:55-136)
# Field order = the nine emits of Octree::encode (src/convert/bytecode.rs:585-596), each traced through its own encoder:
#  [0..2] emit_int(auto_simplify as u8) = 1, emit_int(octree_size) = 8, emit_int(brick_dim) = 2            (:586-588)
#  [3] ObjectPool (object_pool.rs:97-108) = l first_available l ITEM* e e; ITEM = l reserved content e (:25-36).
#      Three pushes, keys 0, 1, 2; `allocate` (:178-202) bumps first_available only while the next slot exists, so it stops
#      at 2. All three are reserved (1).
#      key 0, the root: NodeContent::Internal(occupied_bits) -> l "##" bits e (bytecode.rs:214-217). The root's 4x4x4 bitmap
#        has cells of 2 voxels: (0,0,0) and (1,1,1) both fall into cell (0,0,0) = bit 0; the 4^3 fill at (4..8)^3 covers cells
#        x,y,z in {2,3}: bit x + 4y + 16z (BITMAP_INDEX_LUT, lut.rs:210-235) = {42,43,46,47,58,59,62,63}
#        -> 0xCC00CC0000000001 = 14699973484109365249
#      key 1, the child at root octant 0 (covers (0..4)^3; 4 = 2 * brick_dim, so it is a Leaf, insert.rs:130-204):
#        NodeContent::Leaf -> l "###" brick*8 e (:218-228). Both voxels are in its octant 0 (the 2^3 brick at (0..2)^3):
#        BrickData::Parted -> l "##b#" len voxel*len "#" e (:111-119), 8 voxels in flat_projection order x + 2y + 4z
#        (math/mod.rs:35-37): voxel 0 = colour 0, no data = 0xFFFF0000 = 4294901760 (pix_visual, node.rs:354-403);
#        voxel 7 = (1,1,1) = colour 1 | data 0 << 16 = 1; the six others empty_marker() = 0xFFFFFFFF. Octants 1..7:
#        BrickData::Empty -> "#b" (:106)
#      key 2, the child at root octant 7 (x >= 4: 1, z >= 4: 2, y >= 4: 4): insert_at_lod fills a whole 4^3 node with one value
#        -> NodeContent::UniformLeaf(Solid(colour 1, no data = 0xFFFF0001 = 4294901761)) -> l "##u#" l "#b#" value e e (:229-232, :107-110)
#  [4] Vec<NodeChildren<u32>> parallel to the nodes (:354-372): root = l "##c##" key*8 e with key 1 in octant 0, key 2 in
#      octant 7, empty_marker() = 4294967295 elsewhere; the two leaves = l "##b##" bitmap e: node 1's 4x4x4 cells are single
#      voxels, (0,0,0) -> bit 0 and (1,1,1) -> bit 1 + 4 + 16 = 21 -> 0x200001 = 2097153; node 2 is full -> u64::MAX
#  [5] Vec<BrickData> node_mips, one per key (types.rs:186): MIP maps are off, all BrickData::Empty -> "#b"
#  [6] colour palette = l (l r g b a e)* e (:32-38) in insertion order; [7] data palette = l 7 e
#  [8] MIPMapStrategy::default() (:439-453; same bytes as in ONE_VOXEL above)
TWO_LEVELS = (
    b"l"
    b"i1e" b"i8e" b"i2e"
    b"l" b"i2e"
    b"l"
    b"l" b"i1e" b"l2:##i14699973484109365249ee" b"e"
    b"l" b"i1e" b"l3:###"
    b"l4:##b#i8e" b"i4294901760e" + b"i4294967295e" * 6 + b"i1e" b"1:#e"
    + b"2:#b" * 7 + b"e" b"e"
    b"l" b"i1e" b"l4:##u#" b"l3:#b#i4294901761ee" b"e" b"e"
    b"e" b"e"
    b"l"
    b"l5:##c##" b"i1e" + b"i4294967295e" * 6 + b"i2e" b"e"
    b"l5:##b##i2097153ee"
    b"l5:##b##i18446744073709551615ee"
    b"e"
    b"l2:#b2:#b2:#be"
    b"l" b"li255ei0ei0ei255ee" b"li0ei255ei0ei255ee" b"e"
    b"li7ee"
    b"l" b"i0e" b"i4e" b"i1ei1e" b"i2ei0e" b"i3ei0e" b"i4ei0e" b"i3e" b"i2ei100e" b"i3ei50e" b"i4ei20e" b"e"
    b"e"
)


def build_two_levels(IO):
    t = IO.new(8, 2)
    assert t.insert((0, 0, 0), (255, 0, 0, 255)) == O.OK
    assert t.insert((1, 1, 1), (0, 255, 0, 255), 7) == O.OK
    assert t.insert_at_lod((4, 4, 4), 4, (0, 255, 0, 255)) == O.OK
    return t


def test_two_level_tree_matches_the_hand_assembled_bytes(IO):
    t = build_two_levels(IO)
    assert IO.to_bytes(t) == TWO_LEVELS
    u = IO.from_bytes(TWO_LEVELS)
    assert u.get((0, 0, 0)) == K((255, 0, 0, 255)) and u.get((1, 1, 1)) == K((0, 255, 0, 255), 7)
    assert u.get((1, 0, 0)) == K() and u.get((3, 3, 3)) == K() and u.get((4, 0, 4)) == K()
    for p in ((4, 4, 4), (7, 7, 7), (5, 6, 4)):
        assert u.get(p) == K((0, 255, 0, 255))
    assert u.structure_hash() == t.structure_hash()
    # with MIP maps switched on afterwards the container changes in exactly two places: the `enabled` flag of the strategy
    # ([8], first integer) and node_mips ([5]), which now holds Parted MIP bricks for the root and its first child (their
    # voxel CONTENT is the business of the reference's mipmap KATs, tests/test_mipmap.py) and stays Empty for the UniformLeaf
    t.switch_albedo_mip_maps(True) if hasattr(t, "switch_albedo_mip_maps") else t.tree.albedo_mip_map_resampling_strategy().switch_albedo_mip_maps(True)
    doc, plain = bdec(IO.to_bytes(t)), bdec(TWO_LEVELS)
    assert doc[8] == [1] + plain[8][1:] and doc[:5] == plain[:5] and doc[6:8] == plain[6:8]
    assert [m[0] if isinstance(m, list) else m for m in doc[5]] == [b"##b#", b"##b#", b"#b"]
    assert all(len(m) == 2 + 8 + 1 and m[1] == 8 and m[-1] == b"#" for m in doc[5][:2])


# ---- (2) a third encoder, from the grammar -----------------------------------------------------------------------
def benc(v) -> bytes:
    if isinstance(v, bool):
        return b"i%de" % int(v)
    if isinstance(v, int):
        return b"i%de" % v
    if isinstance(v, (bytes, str)):
        v = v.encode() if isinstance(v, str) else v
        return b"%d:%s" % (len(v), v)
    return b"l" + b"".join(benc(x) for x in v) + b"e"


def bdec(data: bytes):
    """bencode -> nested lists / ints / bytes"""
    def item(i):
        c = data[i:i + 1]
        if c == b"i":
            j = data.index(b"e", i)
            return int(data[i + 1:j]), j + 1
        if c == b"l":
            out, i = [], i + 1
            while data[i:i + 1] != b"e":
                v, i = item(i)
                out.append(v)
            return out, i + 1
        j = data.index(b":", i)
        n = int(data[i:j])
        return data[j + 1:j + 1 + n], j + 1 + n

    v, end = item(0)
    assert end == len(data)
    return v


def canonical(data: bytes) -> bytes:
    """ObjectPool::free keeps the stale item of a slot it un-reserves (object_pool.rs:213-221), so the reference - and
    the oracle, which follows it literally - writes that dead content out. The product releases a freed node's bricks
    and writes NodeContent::Nothing for the slot. Nothing reads an unreserved item (push overwrites it,
    object_pool.rs:172-176), so the two files describe the same tree; this blanks unreserved slots for comparison."""
    doc = bdec(data)
    for it in doc[3][1]:
        if it[0] == 0:
            it[1] = b"#"
    return benc(doc)


def brick(b):
    if b is None:
        return "#b"                                     # BrickData::Empty, bytecode.rs:72
    if isinstance(b, int):
        return ["#b#", b]                               # Solid, :73-76
    return ["##b#", len(b), *b, "#"]                    # Parted, :77-88


def tree_doc(auto_simplify, size, dim, first_available, nodes, children, colors, datas):
    """nodes: list of (reserved, content); content None | ('I', bits) | ('L', [8 bricks]) | ('U', brick)."""
    def content(c):
        if c is None:
            return "#"
        if c[0] == "I":
            return ["##", c[1]]
        if c[0] == "L":
            return ["###", *[brick(b) for b in c[1]]]
        return ["##u#", brick(c[1])]

    def link(c):
        if c is None:
            return "##x##"
        if isinstance(c, int):
            return ["##b##", c]
        return ["##c##", *c]

    strategy = [0, 4, 1, 1, 2, 0, 3, 0, 4, 0, 3, 2, 100, 3, 50, 4, 20]
    return [int(auto_simplify), size, dim, [first_available, [[int(r), content(c)] for r, c in nodes]],
            [link(c) for c in children], ["#b"] * len(nodes), [list(c) for c in colors], list(datas), strategy]


def test_third_encoder_reproduces_the_literal():
    doc = tree_doc(True, 4, 1, 1,
                   [(1, ("I", 1 << 57)), (1, ("L", [None, None, None, 0xFFFF0000, None, None, None, None]))],
                   [[NIL] * 6 + [1, NIL], 0x00CC00CC00000000], [(10, 20, 30, 255)], [])
    assert benc(doc) == ONE_VOXEL


def test_hand_built_documents_load(IO):
    """A file nobody's encoder wrote: parted and solid bricks, a uniform leaf, user data, an unreserved pool slot,
    the strategy maps in another order, MIP bricks and a non-default MIP strategy."""
    dim, vol = 2, 8
    parted = [0xFFFF0000, NIL, NIL, NIL, NIL, NIL, NIL, 0x00000001]          # colour 0 at (0,0,0); colour 1 + data 0 at (1,1,1)
    nodes = [(1, ("I", 0xFFFFFFFFFFFFFFFF)), (1, ("L", [parted, 0x0000FFFF, None, None, None, None, None, None])),
             (0, None), (1, ("U", 0xFFFF0001))]
    children = [[1, NIL, NIL, NIL, NIL, NIL, NIL, 3], 0xFF, None, 0xFFFFFFFFFFFFFFFF]
    doc = tree_doc(False, 8, dim, 2, nodes, children, [(255, 0, 0, 255), (0, 255, 0, 255)], [7])
    doc[5] = ["#b", ["#b#", 0xFFFF0001], "#b", ["##b#", vol, *([0xFFFF0000] * vol), "#"]]   # node_mips with content
    doc[8] = [1, 2, 4, 0, 1, 1003, 1, 3, 50]                                  # enabled, PosterizeBD(0.0) at level 1, one threshold
    t = IO.from_bytes(benc(doc))
    assert t.get((0, 0, 0)) == K((255, 0, 0, 255))
    assert t.get((1, 1, 1)) == K((0, 255, 0, 255), 7)
    assert t.get((1, 0, 0)) == K()
    assert t.get((2, 0, 0)) == K(data=7)                                      # Solid(no colour | data 0) fills octant 1 of node 1
    assert t.get((3, 1, 1)) == K(data=7)
    assert t.get((0, 0, 2)) == K()
    for p in ((4, 4, 4), (7, 7, 7), (5, 6, 4)):                               # UniformLeaf(Solid(colour 1))
        assert t.get(p) == K((0, 255, 0, 255))
    assert t.get((4, 0, 0)) == K()
    # the MIP bricks and the MIP strategy are part of the tree (bytecode.rs:591-594, :638-668)
    assert t.mip_enabled()
    assert t.get_method_at(1) == (4, 0.0) and t.get_method_at(4)[0] == 0 and t.get_method_at(2)[0] == 0
    assert t.get_new_color_similarity_at(3) == float(np.float32(50) / np.float32(1000)) and t.get_new_color_similarity_at(2) == 0.0
    assert t.sample_root_mip(0, (1, 0, 1)) == K((0, 255, 0, 255))             # node 1: Solid MIP, colour 1
    assert t.sample_root_mip(7, (1, 1, 1)) == K((255, 0, 0, 255))             # node 3: Parted MIP, colour 0 everywhere
    assert t.sample_root_mip(8, (0, 0, 0)) == K()                             # root: no MIP
    # saved again, everything survives
    again = IO.from_bytes(IO.to_bytes(t))
    assert again.structure_hash() == t.structure_hash() and again.mip_hash() == t.mip_hash()
    assert IO.to_bytes(again) == IO.to_bytes(t)
    # the product rejects a MIP voxel pointing beyond the palette like any other voxel (the reference panics on use)
    if IO is ProductIO:
        doc[5][1] = ["#b#", 5]
        with pytest.raises(ValueError):
            IO.from_bytes(benc(doc))


def test_posterize_codes_decode_like_the_reference(IO):
    """MIPResamplingMethods (bytecode.rs:519-569): 3 + thr*1000 / 1003 + thr*1000 with the decoder's exclusive ranges:
    1002 and anything from 2001 up are errors, 1003 (what Posterize(1.0) writes) reads back as PosterizeBD(0.0)."""
    doc = tree_doc(True, 4, 1, 1, [(1, None)], [None], [], [])
    for code, want in [(0, (0, 0.0)), (1, (1, 0.0)), (2, (2, 0.0)), (3, (3, 0.0)), (253, (3, 0.25)), (1001, (3, float(np.float32(998) / np.float32(1000)))),
                       (1003, (4, 0.0)), (1503, (4, 0.5)), (2000, (4, float(np.float32(997) / np.float32(1000))))]:
        doc[8] = [0, 1, 6, code, 0]
        assert IO.from_bytes(benc(doc)).get_method_at(6) == want, code
    for code in (1002, 2001, 5000):
        doc[8] = [0, 1, 6, code, 0]
        with pytest.raises(ValueError):
            IO.from_bytes(benc(doc))


@pytest.mark.parametrize("dim", [1, 2, 4])
def test_trees_with_mips_round_trip_and_both_codecs_agree(dim):
    rng = np.random.default_rng(11 + dim)
    size = 16 * dim
    prod, ora = ProductOctree(size, dim), OracleOctree(size, dim)
    for t in (prod, ora):
        t.switch_albedo_mip_maps(True).set_method_at(2, 3, 0.25).set_method_at(3, 2).set_color_similarity_thr_at(1, 0.03)
    for _ in range(200):
        pos = tuple(int(v) for v in rng.integers(0, size, 3))
        col = (int(rng.integers(1, 5)) * 50, int(rng.integers(0, 3)) * 100, 200, 255)
        clear = rng.integers(0, 6) == 0
        for t in (prod, ora):
            if clear:
                t.clear(pos)
            else:
                t.insert(pos, col)
    bp, bo = prod.tree.to_bytes(), ora.to_bytes()
    assert bp == canonical(bo)
    p2, o2 = ProductIO.from_bytes(bo), OracleOctree.from_bytes(bp)
    for t in (p2, o2):
        assert t.mip_enabled() and t.get_method_at(2) == (3, 0.25) and t.get_method_at(3) == (2, 0.0)
        assert t.structure_hash() == prod.structure_hash() and t.mip_hash() == ora.mip_hash()
    assert p2.tree.to_bytes() == bp and o2.to_bytes() == bp
    # the loaded copies keep updating their MIPs like the originals
    for _ in range(50):
        pos = tuple(int(v) for v in rng.integers(0, size, 3))
        for t in (prod, ora, p2, o2):
            t.insert(pos, (9, 9, 9, 255))
    assert prod.mip_hash() == ora.mip_hash() == p2.mip_hash() == o2.mip_hash()


# ---- (3) the two codecs against each other -----------------------------------------------------------------------
def scene_trees():
    yield scenes.cpu_render_scene()
    yield scenes.dot_cube_scene(64, 8)
    yield scenes.colonnade_scene(128, 4)
    yield scenes.terrain_scene(64, 8, 4321, 1, shell=4)
    yield scenes.criterion_scene(64, 8, 20)


@pytest.mark.parametrize("scene", list(scene_trees()), ids=lambda s: s.name)
def test_product_and_oracle_write_identical_bytes_and_read_each_other(scene):
    prod = scenes.build_tree(scene, S.Octree)
    ora = scenes.build_tree(scene, OracleOctree)
    bp, bo = prod.to_bytes(), ora.to_bytes()
    assert bp == canonical(bo) and bp == canonical(bp)
    p2, o2 = S.Octree.from_bytes(bo), OracleOctree.from_bytes(bp)
    assert p2.structure_hash() == prod.structure_hash() == o2.structure_hash()
    assert p2.to_bytes() == bp and o2.to_bytes() == bp
    n = min(scene.tree_size, 24)
    a, b = prod.get_sweep((0, 0, 0), (n, n, n)), p2.get_sweep((0, 0, 0), (n, n, n))
    assert np.array_equal(a, b)


def test_integers_of_every_length_are_written_like_the_oracle_writes_them():
    """The product's encoder formats integers itself (two-digit table, a cache of recent values): user data of 1 to
    10 digits at every power-of-ten boundary, enough distinct values to evict cache slots, repeated values, and 64-bit
    occupancy masks must come out exactly as the oracle's `std::to_string` writes them."""
    values = sorted({10 ** k + d for k in range(10) for d in (-1, 0, 1) if 0 < 10 ** k + d < 2 ** 32 - 1} | {1, 7, 42, 65535, 65536, 2 ** 31, 2 ** 32 - 2})
    rng = np.random.default_rng(11)
    values += [int(v) for v in rng.integers(1, 2 ** 32 - 1, 3000)]
    prod, ora = S.Octree(64, 4), OracleOctree(64, 4)
    pos = [(x, y, z) for x in range(0, 64, 2) for y in range(0, 64, 8) for z in range(0, 64, 4)]
    for (x, y, z), v in zip(pos, values):
        color = (x * 4 % 256, y * 4 % 256, (v % 251) + 1, 255) if v % 3 else None
        for t in (prod, ora):
            t.insert((x, y, z), color, v)
            t.insert((x + 1, y, z), color, values[(v * 7) % 40])   # repeats of a few values
    bp, bo = prod.to_bytes(), ora.to_bytes()
    assert bp == canonical(bo)
    assert S.Octree.from_bytes(bp).to_bytes() == bp and OracleOctree.from_bytes(bp).to_bytes() == bp


def test_loaded_trees_keep_building_like_the_original():
    """ObjectPool state (reserved flags, first_available) and the palette lookup maps survive the round trip: the same
    edits applied to the original and to the loaded copy give identical bytes (i.e. identical node keys too)."""
    rng = np.random.default_rng(5)
    for dim in (1, 2, 4):
        size = 16 * dim
        prod, ora = S.Octree(size, dim), OracleOctree(size, dim)

        def edit(trees, k):
            for _ in range(k):
                pos = tuple(int(v) for v in rng.integers(0, size, 3))
                op = rng.integers(0, 5)
                col = (int(rng.integers(1, 4)) * 60, 10, 200, 255)
                lod, data = int(2 ** rng.integers(0, 3)) * dim, int(rng.integers(1, 5))
                for t in trees:
                    if op <= 1:
                        t.insert(pos, col)
                    elif op == 2:
                        t.insert_at_lod(pos, lod, col)
                    elif op == 3:
                        t.clear(pos)
                    else:
                        t.insert(pos, col, data)

        edit([prod, ora], 150)
        assert prod.to_bytes() == canonical(ora.to_bytes())
        prod2, ora2 = S.Octree.from_bytes(prod.to_bytes()), OracleOctree.from_bytes(ora.to_bytes())
        edit([prod, ora, prod2, ora2], 150)
        assert prod.to_bytes() == prod2.to_bytes() == canonical(ora.to_bytes()) == canonical(ora2.to_bytes())


# ---- the reference's own tests, src/convert/bytecode_tests.rs ------------------------------------------------------
def test_octree_file_io(tmp_path):  # :182-222, product only (save / load are file wrappers around to_bytes / from_bytes)
    red = 0xFF0000FF
    tree = S.Octree(8, 1)
    tree.insert_at_lod((0, 0, 0), 4, red)
    tree.clear_at_lod((0, 0, 0), 2)
    path = tmp_path / "test_junk_octree"
    tree.save(str(path))
    copy = S.Octree.load(str(path))
    hits = 0
    for x in range(4):
        for y in range(4):
            for z in range(4):
                assert tree.get((x, y, z)) == copy.get((x, y, z))
                if tree.get((x, y, z)).is_some():
                    assert tree.get((x, y, z)).albedo == S.Albedo(255, 0, 0, 255)
                    hits += 1
    assert hits == 64 - 8
    assert path.read_bytes() == tree.to_bytes()
    with pytest.raises(S.OctreeError) as e:
        S.Octree.load(str(tmp_path / "missing"))
    assert e.value.code == S.api.E_IO
    with pytest.raises(S.OctreeError) as e:
        tree.save(str(tmp_path / "no_such_dir" / "x"))
    assert e.value.code == S.api.E_IO


def fill_and_check(IO, size, dim, lo, color_of):
    t = IO.new(size, dim)
    for x in range(lo, size):
        for y in range(lo, size):
            for z in range(lo, size):
                assert t.insert((x, y, z), color_of(x, y, z)) == O.OK
    u = IO.from_bytes(IO.to_bytes(t))
    for x in range(lo, size):
        for y in range(lo, size):
            for z in range(lo, size):
                assert u.get((x, y, z)) == K(color_of(x, y, z)), (x, y, z)
    assert u.structure_hash() == t.structure_hash()


def test_big_octree_serialize(IO):  # :225-250 (fill range shortened from 28^3 to 12^3 voxels to keep the CPU suite fast)
    fill_and_check(IO, 128, 1, 116, lambda x, y, z: x + y + z)


def test_small_octree_serialize_where_dim_is_1(IO):  # :253-267
    t = IO.new(2, 1)
    t.insert((0, 0, 0), 1)
    assert IO.from_bytes(IO.to_bytes(t)).get((0, 0, 0)) == K(1)


def test_octree_serialize_where_dim_is_1(IO):  # :270-301
    fill_and_check(IO, 4, 1, 0, lambda x, y, z: (x << 24) + (y << 16) + (z << 8) + 0xFF)


def test_octree_serialize_where_dim_is_2(IO):  # :304-334
    fill_and_check(IO, 4, 2, 0, lambda x, y, z: (x << 24) + (y << 16) + (z << 8) + 0xFF)


def test_big_octree_serialize_where_dim_is_2(IO):  # :337-364
    fill_and_check(IO, 128, 2, 116, lambda x, y, z: (x << 24) + (y << 16) + (z << 8) + 0xFF)


# ---- malformed input -------------------------------------------------------------------------------------------
def test_malformed_bytes_are_rejected_not_crashed(IO):
    good = ONE_VOXEL
    bad_inputs = [b"", b"le", b"i1e", b"l" + good, good + b"e", good[:-1], good[: len(good) // 2],
                  good.replace(b"3:###", b"3:#?#"), good.replace(b"5:##c##", b"5:##q##"),
                  good.replace(b"i1ei4ei1e", b"i1ei4ei3e", 1),   # brick_dim 3: InvalidBrickDimension, like Octree::new
                  good.replace(b"i4294901760e", b"i-5e"), good.replace(b"2:#b", b"9:#b", 1)]
    for i, b in enumerate(bad_inputs):
        with pytest.raises(ValueError):
            IO.from_bytes(b)
    rng = np.random.default_rng(11)
    for _ in range(300):  # random corruption must never crash; it may still decode to some tree
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(32, 127))
        try:
            t = IO.from_bytes(bytes(b))
            t.get((1, 2, 3))
        except ValueError:
            pass


def test_product_rejects_trees_it_could_not_traverse_safely():
    """The reference would bounds-panic on a palette index beyond the palette and never return from a child cycle; the
    product refuses such files (get(), the GPU serialiser and the kernels trust a loaded tree)."""
    def doc(**kw):
        a = dict(nodes=[(1, ("I", 1 << 57)), (1, ("L", [None, None, None, 0xFFFF0000, None, None, None, None]))],
                 children=[[NIL] * 6 + [1, NIL], 0x00CC00CC00000000], colors=[(10, 20, 30, 255)], datas=[])
        a.update(kw)
        return benc(tree_doc(True, 4, 1, 1, a["nodes"], a["children"], a["colors"], a["datas"]))

    assert S.Octree.from_bytes(doc()).get((1, 2, 3)).albedo == S.Albedo(10, 20, 30, 255)
    bad = [
        doc(colors=[]),                                                                      # colour index 0 without a palette entry
        doc(nodes=[(1, ("I", 1)), (1, ("L", [0x0003FFFF] + [None] * 7))]),                   # data index 3, empty data palette
        doc(children=[[NIL] * 6 + [0, NIL], 0]),                                             # the root is its own child
        doc(nodes=[(1, ("I", 1)), (1, ("I", 1))], children=[[1] + [NIL] * 7, [1] + [NIL] * 7]),   # node 1 -> node 1
        doc(nodes=[(1, ("I", 1)), (1, ("I", 1)), (1, ("I", 1))],
            children=[[1, 1] + [NIL] * 6, [2] + [NIL] * 7, None]),                           # one child under two octants
        doc(nodes=[(1, ("I", 1)), (1, ("I", 1)), (1, ("I", 1)), (1, None)],
            children=[[1] + [NIL] * 7, [2] + [NIL] * 7, [3] + [NIL] * 7, None]),             # deeper than a 4/1 tree can be
    ]
    for i, b in enumerate(bad):
        with pytest.raises(S.OctreeError) as e:
            S.Octree.from_bytes(b)
        assert e.value.code == S.api.E_DECODE, i
