"""The N > 1 path on CPU: world_size 2 and 3 over gloo.  Each rank fills the tiled buffer it would render (a hash of the
global pixel coordinate, standing in for the rgba8 value), the buffers are gathered on rank 0 exactly as bench.py
gathers them over NCCL, and the frame is laid out with the host mirror of rtb_untile.  Checks the partition covers every
pixel once, the equal-count gather, and the slot <-> pixel mapping (which the GPU test compares with the CUDA side)."""
import os
import socket

import numpy as np
import pytest

from igx_raytracing_b200 import tiles


def pixel_value(x, y):
    return ((x.astype(np.uint64) * 73856093) ^ (y.astype(np.uint64) * 19349663) ^ 0xA5A5).astype(np.uint32)


def _worker(rank, world, port, w, h, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slots = tiles.slots_per_rank(w, h, world)
    x, y, valid = tiles.slot_pixels(w, h, rank, world)
    local = np.zeros(slots, np.uint32)
    local[valid] = pixel_value(x[valid], y[valid])
    t = torch.from_numpy(local.view(np.int32).copy())
    gathered = [torch.zeros(slots, dtype=torch.int32) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    # max-over-ranks timing plumbing used by bench.py
    ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert ms.item() == float(world)
    if rank == 0:
        g = np.stack([v.numpy().view(np.uint32) for v in gathered])
        np.save(out_path, tiles.untile(g, w, h, world))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,w,h", [(2, 333, 200), (3, 640, 360), (2, 31, 17)])
def test_partition_gather_untile(tmp_path, world, w, h):
    import torch.multiprocessing as mp
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, out), nprocs=world, join=True)
    got = np.load(out)
    yy, xx = np.mgrid[0:h, 0:w]
    assert np.array_equal(got, pixel_value(xx.reshape(-1), yy.reshape(-1)).reshape(h, w))


def test_partition_covers_every_pixel_once():
    for (w, h, n) in [(1920, 1080, 8), (3840, 2160, 4), (640, 360, 3), (33, 65, 2), (1, 1, 8)]:
        count = np.zeros((h, w), np.int32)
        per_rank = []
        for r in range(n):
            x, y, valid = tiles.slot_pixels(w, h, r, n)
            np.add.at(count, (y[valid], x[valid]), 1)
            per_rank.append(int(valid.sum()))
            assert tiles.local_blocks(w, h, r, n) * 1024 <= tiles.slots_per_rank(w, h, n)
        assert (count == 1).all()
        if w * h > 100000:   # interleaved blocks balance the ranks
            assert max(per_rank) - min(per_rank) <= 2 * 1024
