"""Adapter giving the product's Octree (shocovox_b200.api.Octree, over the C ABI) the same duck-typed surface as
tests/oracle_lib.OracleOctree, so one test body runs against both."""
import numpy as np

import oracle_lib as O
import shocovox_b200 as S


class ProductOctree:
    def __init__(self, size, brick_dim):
        try:
            self.tree = S.Octree(size, brick_dim)
        except S.OctreeError as e:
            raise ValueError(e.code)

    def set_auto_simplify(self, v):
        self.tree.set_auto_simplify(v)

    def _call(self, fn, *a, **k):
        try:
            fn(*a, **k)
            return O.OK
        except S.OctreeError as e:
            return e.code

    def insert(self, pos, albedo=None, data=None):
        return self._call(self.tree.insert, pos, albedo, data)

    def update(self, pos, albedo=None, data=None):
        return self._call(self.tree.update, pos, albedo, data)

    def insert_at_lod(self, pos, size, albedo=None, data=None):
        return self._call(self.tree.insert_at_lod, pos, size, albedo, data)

    def clear(self, pos):
        return self._call(self.tree.clear, pos)

    def clear_at_lod(self, pos, size):
        return self._call(self.tree.clear_at_lod, pos, size)

    def insert_batch(self, xyz, rgba, lod=None):
        return self._call(self.tree.insert_batch, xyz, rgba, lod)

    def get(self, pos):
        e = self.tree.get(pos)
        if e.kind == S.api.ENTRY_EMPTY:
            return (O.EMPTY,)
        rgba = (e.albedo.r, e.albedo.g, e.albedo.b, e.albedo.a) if e.albedo is not None else None
        if e.kind == S.api.ENTRY_VISUAL:
            return (O.VISUAL, rgba)
        if e.kind == S.api.ENTRY_INFORMATIVE:
            return (O.INFORMATIVE, e.data)
        return (O.COMPLEX, rgba, e.data)

    def get_sweep(self, origin, extent):
        return self.tree.get_sweep(origin, extent)

    def structure_hash(self):
        return self.tree.structure_hash()

    def node_count(self):
        return self.tree.node_count()
