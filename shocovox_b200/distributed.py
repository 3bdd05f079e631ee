"""Multi-GPU orchestration of the primary-ray path: one process per GPU, torch.distributed as plumbing.

Rays are independent and the tree is read-only while rendering, so the path shards without a data-path exchange:

  * pose mode  - camera pose k of a batch goes to rank k % world (BASELINE config 5); every rank keeps a full replica
                 of the tree; nothing is exchanged while rendering. Weak scaling.
  * tile mode  - ONE frame is split into bands of `band` image rows, band b goes to rank b % world (interleaved for
                 load balance: sky vs geometry; BASELINE config 4). The bands have to end up in one framebuffer:
                   - `gather_bands`: every rank renders into a compact band-major buffer and the buffers are
                     all-gathered (NCCL over NVLink; gloo in the CPU tests) and de-interleaved;
                   - fused (the product path, csrc/multi_gpu.cu): the traversal kernel of rank r stores its pixels
                     straight into rank 0's framebuffer through a CUDA-IPC peer mapping and the hand-over runs on
                     device-side flags (api.OctreeGPUView.gather_open / gather_join); `open_gather` below does the one-off
                     handle exchange, after which a frame needs no host barrier and no collective at all.

The functions here are backend-agnostic (they take tensors and a process group), which is what lets the world_size-2
gloo tests in tests/test_distributed_cpu.py cover the host logic without a GPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_rows(height: int, rank: int, world: int, band: int) -> List[int]:
    """Image rows rendered by `rank`: rows r with (r // band) % world == rank (svx_view_set_shard)."""
    return [r for r in range(height) if (r // band) % world == rank]


def local_row_count(height: int, rank: int, world: int, band: int) -> int:
    """Rows of the compact (band-major) buffer of `rank`: owned bands * band (a partial last band is padded)."""
    bands = (height + band - 1) // band
    owned = len(range(rank, bands, world))
    return owned * band


def padded_local_rows(height: int, world: int, band: int) -> int:
    """Equal-sized compact buffers for collectives: the largest local_row_count over the ranks."""
    return max(local_row_count(height, r, world, band) for r in range(world))


def poses_for_rank(n_poses: int, rank: int, world: int) -> List[int]:
    """Pose k of a batch is rendered by rank k % world."""
    return list(range(rank, n_poses, world))


def deinterleave(gathered, height: int, world: int, band: int):
    """[world, local_rows, W] band-major compact buffers -> [height, W] image (torch tensor in, torch tensor out)."""
    import torch

    w = gathered.shape[-1]
    local_rows = gathered.shape[1]
    nb_local = local_rows // band
    # gathered[r, j*band + i] is image row (j*world + r)*band + i
    g = gathered.reshape(world, nb_local, band, w).permute(1, 0, 2, 3).reshape(nb_local * world * band, w)
    return g[:height].contiguous() if isinstance(g, torch.Tensor) else g[:height]


def gather_bands(local, height: int, world: int, band: int, group=None):
    """All-gathers the compact band-major buffers of every rank and returns the assembled [height, W] image.

    `local` is [padded_local_rows, W] on any device (CUDA with the nccl backend, CPU with gloo)."""
    import torch
    import torch.distributed as dist

    flat = local.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return deinterleave(out.reshape((world,) + tuple(local.shape)), height, world, band)


def device_tensor(ptr: int, shape: Sequence[int], dtype_str: str, device_index: int):
    """Wraps a raw CUDA device pointer (e.g. a view's framebuffer plane) as a torch tensor, without copying."""
    import numpy as np
    import torch

    class _Wrapper:
        def __init__(self):
            self.__cuda_array_interface__ = {
                "shape": tuple(int(s) for s in shape), "typestr": np.dtype(dtype_str).str, "data": (int(ptr), False),
                "version": 3, "strides": None,
            }

    with torch.cuda.device(device_index):
        return torch.as_tensor(_Wrapper(), device=f"cuda:{device_index}")


def broadcast_bytes(blob, src_rank: int = 0, group=None) -> bytes:
    """Broadcasts a small byte string (e.g. the 128-byte gather handle of rank `src_rank`'s root view) to every rank."""
    import torch.distributed as dist

    box = [blob if dist.get_rank(group) == src_rank else None]
    dist.broadcast_object_list(box, src=src_rank, group=group)
    return box[0]


def open_gather(view, rank: int, world: int, band: int = 8, wire: int = 0, group=None):
    """One-off set-up of the fused tile-sharded gather over `world` processes (one per GPU): rank 0's view becomes the
    root whose framebuffer assembles the frame, every other rank's view joins with the broadcast handle. Afterwards each
    rank just calls view.render(): the synchronisation per frame is on the devices (svx_view_gather_*)."""
    import torch.distributed as dist

    handle = view.gather_open(world, band, wire) if rank == 0 else None
    handle = broadcast_bytes(handle, 0, group)
    if rank != 0:
        view.gather_join(rank, handle)
    dist.barrier(group)  # every peer has mapped the root's frame before the first frame is rendered
    return handle


class SharedPinnedPlanes:
    """`sets` x (hit_id u32, albedo u32, distance f32) full-frame host planes in ONE POSIX shared-memory file that every
    rank maps and page-locks (cudaHostRegister), so that each rank's GPU copies the rows it rendered straight into the
    common frame over its own PCIe link (svx_view_render_to_host on a sharded view). Rank 0 creates and unlinks the file."""

    def __init__(self, tag: str, n_px: int, sets: int, rank: int, group=None):
        import numpy as np
        import torch
        import torch.distributed as dist

        self.path = f"/dev/shm/svx_planes_{tag}"
        self.rank = rank
        nbytes = sets * 3 * n_px * 4
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        dist.barrier(group)
        self.map = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(nbytes,))
        self.ptr = self.map.ctypes.data
        self.nbytes = nbytes
        self.map[:] = 0  # touch every page before pinning
        rc = torch.cuda.cudart().cudaHostRegister(self.ptr, nbytes, 0)
        self.pinned = int(rc) == 0
        self.sets = []
        for s in range(sets):
            base = s * 3 * n_px * 4
            self.sets.append([self.ptr + base, self.ptr + base + n_px * 4, self.ptr + base + 2 * n_px * 4])
        self.n_px = n_px
        dist.barrier(group)

    def plane(self, s: int, k: int, dtype):
        import numpy as np

        off = (s * 3 + k) * self.n_px * 4
        return self.map[off:off + self.n_px * 4].view(dtype)

    def close(self, group=None):
        import os
        import torch
        import torch.distributed as dist

        if self.pinned:
            torch.cuda.cudart().cudaHostUnregister(self.ptr)
        dist.barrier(group)
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass
