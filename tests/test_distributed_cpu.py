"""Host logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo process groups, no GPU.

What is covered: the row-band sharding map (it must be the one the kernel uses), the compact band-major layout and its
all-gather + de-interleave, pose assignment, and the IPC-handle broadcast helper."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shocovox_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pixel(row, col):
    return (row * 131071 + col * 8191) & 0x7FFFFFFF


def _worker(rank, world, port, height, width, band, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # what the kernel does for this shard: compact band-major rows, padded to the common size
        rows = D.shard_rows(height, rank, world, band)
        local = torch.full((D.padded_local_rows(height, world, band), width), -1, dtype=torch.int64)
        for lr in range(D.local_row_count(height, rank, world, band)):
            b, within = divmod(lr, band)
            row = (b * world + rank) * band + within
            if row < height:
                assert row in rows
                local[lr] = torch.tensor([_pixel(row, c) for c in range(width)])
        image = D.gather_bands(local, height, world, band)
        want = torch.tensor([[_pixel(r, c) for c in range(width)] for r in range(height)])
        ok = bool(torch.equal(image, want))
        blob = D.broadcast_bytes(bytes([rank + 1]) * 128, src_rank=0)  # the gather handle of rank 0's root view
        ok = ok and blob == bytes([1]) * 128
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == float(world)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,height,width,band", [(2, 203, 17, 8), (2, 64, 5, 16), (3, 100, 9, 4)])
def test_gather_bands_assembles_the_frame(world, height, width, band):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, height, width, band, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def test_shard_rows_partition_the_image():
    for height, world, band in [(1080, 8, 8), (2160, 4, 16), (203, 3, 8), (7, 2, 8)]:
        seen = []
        for r in range(world):
            rows = D.shard_rows(height, r, world, band)
            assert len(rows) <= D.local_row_count(height, r, world, band)
            seen += rows
        assert sorted(seen) == list(range(height))


def test_pose_assignment_covers_the_batch():
    for n, world in [(256, 8), (7, 2), (3, 4)]:
        got = sorted(k for r in range(world) for k in D.poses_for_rank(n, r, world))
        assert got == list(range(n))
        assert max(len(D.poses_for_rank(n, r, world)) for r in range(world)) - min(len(D.poses_for_rank(n, r, world)) for r in range(world)) <= 1


def test_deinterleave_matches_kernel_row_map():
    world, band, height, width = 4, 8, 90, 3
    lr = D.padded_local_rows(height, world, band)
    g = torch.zeros((world, lr, width), dtype=torch.int64)
    for r in range(world):
        for j in range(lr):
            b, within = divmod(j, band)
            g[r, j] = (b * world + r) * band + within  # the image row the kernel maps local row j to
    img = D.deinterleave(g, height, world, band)
    assert torch.equal(img[:, 0], torch.arange(height))
