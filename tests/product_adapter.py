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

    # ---- MIP maps (StrategyUpdater surface, duck-typed like tests/oracle_lib.OracleOctree)
    def _strategy(self):
        return self.tree.albedo_mip_map_resampling_strategy()

    def switch_albedo_mip_maps(self, enabled):
        self._strategy().switch_albedo_mip_maps(enabled)
        return self

    def mip_enabled(self):
        return self._strategy().is_enabled()

    def set_method_at(self, level, method, thr=0.0):
        self._strategy().set_method_at(level, method, thr)
        return self

    def get_method_at(self, level):
        return self._strategy().get_method_at(level)

    def set_color_similarity_thr_at(self, level, thr):
        self._strategy().set_color_similarity_thr_at(level, thr)
        return self

    def get_new_color_similarity_at(self, level):
        return self._strategy().get_new_color_similarity_at(level)

    def mip_reset(self):
        self._strategy().reset()
        return self

    def recalculate_mips(self):
        self._strategy().recalculate_mips()
        return self

    def sample_root_mip(self, octant, pos):
        e = self._strategy().sample_root_mip(octant, pos)
        if e.kind == S.api.ENTRY_EMPTY:
            return (O.EMPTY,)
        rgba = (e.albedo.r, e.albedo.g, e.albedo.b, e.albedo.a) if e.albedo is not None else None
        if e.kind == S.api.ENTRY_VISUAL:
            return (O.VISUAL, rgba)
        if e.kind == S.api.ENTRY_INFORMATIVE:
            return (O.INFORMATIVE, e.data)
        return (O.COMPLEX, rgba, e.data)

    def mip_hash(self):
        return self._strategy().mip_hash()
