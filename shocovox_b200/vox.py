"""MagicaVoxel `.vox` import: the input format of the reference's large examples.

Follows the reference's loader (src/convert/magicavoxel.rs, which sits on the third-party `dot_vox` 5.1.1 parser):

  * chunk parser for the published MagicaVoxel format (SIZE / XYZI / RGBA / nTRN / nGRP / nSHP); palette indices are
    stored 1-based in the file and used 0-based (`dot_vox`: "i is 1 less than the value stored in the source file");
  * scene-graph walk of `iterate_vox_tree` (:105-197): translations accumulate, `_r` rotations are
    `parent_rotation * parse_rotation_matrix(byte)` (and reset to identity where a transform has no `_r`, as the
    reference does), shape nodes emit their frame-0 models;
  * placement of `load_vox_data_internal` (:349-385): model origin = translation - size/2 (integer division, with the
    -1 correction for negative half sizes), voxel positions rotated, then Rzup -> Lyup = (x, z, y)
    (src/spatial/math/mod.rs:195-199), relative to the minimum corner found by `load_vox_file_internal` (:297-347);
  * tree size = next power of two of the largest extent (:266-271).

The result is a voxel list in the reference's insertion order for `Octree.insert_batch`.

This module is the independent SECOND reader (numpy) that cross-checks the product's loader, csrc/vox_import.cpp behind
`svx_octree_load_vox` / `Octree.load_vox_file`, plus a small `.vox` writer for tests. Like the C++ loader it refuses files
without an RGBA chunk (dot_vox substitutes MagicaVoxel's built-in default palette, which is not reproduced here) and files
without a scene graph (the reference panics on `vox_tree.scenes[0]`).

Parity status: the parsing layer is UNPINNED (the reference's loader cannot be run here and `dot_vox` is not in the
checkout); the placement arithmetic is pinned by the rotation known-answer test of magicavoxel.rs:392-413, and two
independent implementations agree on the reference's own assets (tests/test_vox_import.py).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass
class VoxModel:
    size: Tuple[int, int, int]
    voxels: np.ndarray  # [n, 4] u8: x, y, z, palette index (0-based)


@dataclass
class VoxScene:
    models: List[VoxModel] = field(default_factory=list)
    palette: Optional[np.ndarray] = None  # the RGBA chunk; None = the file has none
    nodes: Dict[int, dict] = field(default_factory=dict)  # scene graph by node id


def parse_rotation_matrix(b: int) -> np.ndarray:
    """magicavoxel.rs:60-83: bits 0-1 / 2-3 give the column of the non-zero entry of rows 0 / 1, bits 4-6 the signs."""
    m = np.zeros((3, 3), dtype=np.int64)
    c0 = b & 0x3
    c1 = (b >> 2) & 0x3
    c2 = (~(c0 ^ c1)) & 0x3
    m[0, c0] = -1 if b & 0x10 else 1
    m[1, c1] = -1 if b & 0x20 else 1
    m[2, c2] = -1 if b & 0x40 else 1
    return m


def _read_dict(buf: bytes, pos: int):
    (n,) = struct.unpack_from("<i", buf, pos)
    pos += 4
    out = {}
    for _ in range(n):
        (kl,) = struct.unpack_from("<i", buf, pos)
        pos += 4
        k = buf[pos:pos + kl].decode("utf-8", "replace")
        pos += kl
        (vl,) = struct.unpack_from("<i", buf, pos)
        pos += 4
        v = buf[pos:pos + vl].decode("utf-8", "replace")
        pos += vl
        out[k] = v
    return out, pos


def parse_vox(data: bytes) -> VoxScene:
    if data[:4] != b"VOX ":
        raise ValueError("not a MagicaVoxel file")
    scene = VoxScene()
    pos = 8
    if data[pos:pos + 4] != b"MAIN":
        raise ValueError("MAIN chunk missing")
    _, main_children = struct.unpack_from("<ii", data, pos + 4)
    pos += 12
    end = pos + main_children
    pending_size: Optional[Tuple[int, int, int]] = None
    while pos + 12 <= end:
        cid = data[pos:pos + 4]
        n_content, n_children = struct.unpack_from("<ii", data, pos + 4)
        body = pos + 12
        if cid == b"SIZE":
            pending_size = struct.unpack_from("<iii", data, body)
        elif cid == b"XYZI":
            (n,) = struct.unpack_from("<i", data, body)
            v = np.frombuffer(data, dtype=np.uint8, count=4 * n, offset=body + 4).reshape(n, 4).copy()
            v[:, 3] = np.where(v[:, 3] > 0, v[:, 3] - 1, 0)  # 1-based in the file; dot_vox: `i.saturating_sub(1)`
            scene.models.append(VoxModel(tuple(pending_size or (0, 0, 0)), v))
        elif cid == b"RGBA":
            pal = np.frombuffer(data, dtype=np.uint8, count=1024, offset=body).reshape(256, 4).copy()
            scene.palette = pal
        elif cid in (b"nTRN", b"nGRP", b"nSHP"):
            p = body
            (node_id,) = struct.unpack_from("<i", data, p)
            p += 4
            attrs, p = _read_dict(data, p)
            if cid == b"nTRN":
                child, _reserved, layer, n_frames = struct.unpack_from("<iiii", data, p)
                p += 16
                frames = []
                for _ in range(n_frames):
                    f, p = _read_dict(data, p)
                    frames.append(f)
                scene.nodes[node_id] = {"kind": "transform", "child": child, "frames": frames, "attrs": attrs, "layer": layer}
            elif cid == b"nGRP":
                (n,) = struct.unpack_from("<i", data, p)
                p += 4
                children = list(struct.unpack_from(f"<{n}i", data, p)) if n else []
                scene.nodes[node_id] = {"kind": "group", "children": children, "attrs": attrs}
            else:
                (n,) = struct.unpack_from("<i", data, p)
                p += 4
                models = []
                for _ in range(n):
                    (mid,) = struct.unpack_from("<i", data, p)
                    p += 4
                    ma, p = _read_dict(data, p)
                    models.append({"model_id": mid, "attrs": ma})
                scene.nodes[node_id] = {"kind": "shape", "models": models, "attrs": attrs}
        pos = body + n_content + n_children
    return scene


def iterate_models(scene: VoxScene, frame: int = 0):
    """iterate_vox_tree, magicavoxel.rs:105-197: yields (model, translation[3], rotation[3,3])."""
    if 0 not in scene.nodes:
        raise ValueError("no scene graph (the reference panics on vox_tree.scenes[0], magicavoxel.rs:112)")
    root = scene.nodes[0]
    if root["kind"] != "transform":
        raise ValueError("the root node of a MagicaVoxel scene graph should be a transform")
    stack = [[root["child"], np.zeros(3, dtype=np.int64), np.eye(3, dtype=np.int64), 0]]
    while stack:
        node_id, translation, rotation, index = stack[-1]
        node = scene.nodes[node_id]
        if node["kind"] == "transform":
            frames = node["frames"]
            used = frame if frame < len(frames) else 0
            fr = frames[used] if frames else {}
            t = translation + np.array([int(v) for v in fr["_t"].split(" ")], dtype=np.int64) if "_t" in fr else translation
            r = rotation @ parse_rotation_matrix(int(fr["_r"])) if "_r" in fr else np.eye(3, dtype=np.int64)
            if index == 0:
                stack[-1][3] += 1
                stack.append([node["child"], t, r, 0])
            else:
                stack.pop()
        elif node["kind"] == "group":
            if index < len(node["children"]):
                stack[-1][3] += 1
                stack.append([node["children"][index], translation, rotation, 0])
            else:
                stack.pop()
        else:
            for m in node["models"]:
                if int(m["attrs"].get("_f", "0")) == frame:
                    yield scene.models[m["model_id"]], translation, rotation
            stack.pop()
            if stack:
                stack[-1][3] += 1


def _trunc_half(v: np.ndarray) -> np.ndarray:
    """Rust integer division by 2 truncates toward zero."""
    return np.where(v >= 0, v // 2, -((-v) // 2))


def _rzup_to_lyup(v: np.ndarray) -> np.ndarray:
    """convert_coordinate(Rzup -> Lyup) = (x, z, y) (src/spatial/math/mod.rs:195-199)."""
    return np.stack([v[..., 0], v[..., 2], v[..., 1]], axis=-1)


def load_vox(path_or_bytes, brick_dimension: int = 8):
    """-> (tree_size, xyz u32[n,3], rgba u8[n,4]) in the reference's insertion order (Octree::load_vox_file)."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    scene = parse_vox(bytes(data))
    if scene.palette is None:
        raise ValueError("no RGBA chunk: MagicaVoxel's built-in default palette is not reproduced by this reader")
    placed = list(iterate_models(scene, 0))
    if not placed:
        raise ValueError("no models in the file")
    lo = np.full(3, np.iinfo(np.int32).max, dtype=np.int64)
    hi = np.full(3, np.iinfo(np.int32).min, dtype=np.int64)
    for model, t, r in placed:  # load_vox_file_internal, :297-347
        half = _trunc_half(r @ np.array(model.size, dtype=np.int64))
        lo = np.minimum(lo, np.minimum(t - half, t + half))
        hi = np.maximum(hi, np.maximum(t - half, t + half))
    min_lyup, max_lyup = _rzup_to_lyup(lo), _rzup_to_lyup(hi)
    extent = int((max_lyup - min_lyup).max())
    tree_size = 1 << int(np.ceil(np.log2(np.float32(max(extent, 1)))))
    if tree_size < 2 * brick_dimension:
        raise ValueError(f"model extent {extent} gives tree size {tree_size} < 2 * brick dimension (Octree::new fails, the reference panics)")
    min_rzup = _rzup_to_lyup(min_lyup)  # Lyup -> Rzup is the same swap
    xyz, rgba = [], []
    for model, t, r in placed:  # load_vox_data_internal, :349-385
        half = _trunc_half(r @ np.array(model.size, dtype=np.int64))
        bottom_left = t - half - min_rzup + np.where(half < 0, -1, 0)
        v = model.voxels[:, :3].astype(np.int64) @ r.T
        p = _rzup_to_lyup(bottom_left[None, :] + v)
        xyz.append(p)
        rgba.append(scene.palette[model.voxels[:, 3]])
    xyz = np.concatenate(xyz)
    rgba = np.concatenate(rgba)
    inside = ((xyz >= 0) & (xyz < tree_size)).all(axis=1)
    if not inside.all():
        raise ValueError("voxel outside of the computed tree bounds (the reference panics here)")
    return tree_size, xyz.astype(np.uint32), rgba.astype(np.uint8)


def load_vox_file(path_or_bytes, brick_dimension: int = 8, mip_enabled: bool = False, mip_methods=None,
                  mip_color_similarity=None):
    """`Octree::load_vox_file(brick_dimension, path)` and, with the MIP arguments, `MIPMapStrategy::load_vox_file`
    (src/convert/magicavoxel.rs:207-250): the strategy is installed on the empty tree BEFORE the voxels are inserted, so
    every insert refreshes the MIPs as it goes (insert.rs:371) - `MIPMapStrategy::default().set_enabled(true)
    .load_vox_file(..)` is what examples/minecraft.rs:57-60 and examples/sponza.rs:66-67 do.
    mip_methods = {level: method | (method, threshold)}, mip_color_similarity = {level: threshold}: applied on top of
    MIPMapStrategy::default(), like the reference's builder calls."""
    from .api import Octree

    tree_size, xyz, rgba = load_vox(path_or_bytes, brick_dimension)
    tree = Octree(tree_size, brick_dimension)
    strategy = tree.albedo_mip_map_resampling_strategy()
    strategy.set_method(list((mip_methods or {}).items()))
    strategy.set_color_similarity_thr(list((mip_color_similarity or {}).items()))
    if mip_enabled:
        strategy.switch_albedo_mip_maps(True)  # the tree is still empty: nothing to recalculate (mipmap.rs:866-871)
    tree.insert_batch(xyz, rgba)
    return tree


def write_vox(models, palette: Optional[np.ndarray] = None, placements=None) -> bytes:
    """Minimal .vox writer (for tests): models = [(size, voxels[n,4] with 0-based palette index)], optional scene graph
    placements = [(translation(3), rotation_byte or None)] per model."""
    def chunk(cid: bytes, content: bytes, children: bytes = b"") -> bytes:
        return cid + struct.pack("<ii", len(content), len(children)) + content + children

    def wdict(d: dict) -> bytes:
        out = struct.pack("<i", len(d))
        for k, v in d.items():
            kb, vb = k.encode(), str(v).encode()
            out += struct.pack("<i", len(kb)) + kb + struct.pack("<i", len(vb)) + vb
        return out

    body = b""
    for size, vox in models:
        vox = np.asarray(vox, dtype=np.int64)
        stored = vox.copy()
        stored[:, 3] = (stored[:, 3] + 1) & 0xFF
        body += chunk(b"SIZE", struct.pack("<iii", *size))
        body += chunk(b"XYZI", struct.pack("<i", len(vox)) + stored.astype(np.uint8).tobytes())
    if placements is not None:
        n = len(models)
        body += chunk(b"nTRN", struct.pack("<i", 0) + wdict({}) + struct.pack("<iiii", 1, -1, -1, 1) + wdict({}))
        body += chunk(b"nGRP", struct.pack("<i", 1) + wdict({}) + struct.pack(f"<i{n}i", n, *[2 + 2 * i for i in range(n)]))
        for i, (t, rot) in enumerate(placements):
            fr = {"_t": " ".join(str(int(c)) for c in t)}
            if rot is not None:
                fr["_r"] = int(rot)
            body += chunk(b"nTRN", struct.pack("<i", 2 + 2 * i) + wdict({}) + struct.pack("<iiii", 3 + 2 * i, -1, 0, 1) + wdict(fr))
            body += chunk(b"nSHP", struct.pack("<i", 3 + 2 * i) + wdict({}) + struct.pack("<i", 1) + struct.pack("<i", i) + wdict({}))
    if palette is not None:
        body += chunk(b"RGBA", np.asarray(palette, dtype=np.uint8).reshape(256, 4).tobytes())
    return b"VOX " + struct.pack("<i", 150) + chunk(b"MAIN", b"", body)
