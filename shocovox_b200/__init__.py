"""shocovox_b200 — B200-native primary-ray traversal of the shocovox brick-leaf sparse voxel octree.

The package holds only what the hot path needs: the CUDA kernels + C ABI (csrc/, built into libshocovox_b200.so),
the host-side mirror of the reference crate's API (api.py) and the synthetic benchmark scenes (scenes.py).
"""
from .api import (  # noqa: F401
    Albedo, Octree, OctreeEntry, OctreeError, OctreeGPUHost, OctreeGPUView, Ray, RayHit, Viewport,
    GLASS_AT_FOV, GLASS_AT_FRUSTUM_Z, MISS, cuda_device_count, entry, lib, library_path, normalized,
    MultiGPU, WIRE_THREE_PLANES, WIRE_ID_DISTANCE, E_TIMEOUT,
    StrategyUpdater, F32_MAX, MIP_BOX_FILTER, MIP_POINT_FILTER, MIP_POINT_FILTER_BD, MIP_POSTERIZE, MIP_POSTERIZE_BD,
)

__version__ = "0.1.0"
