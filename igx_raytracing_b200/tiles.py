"""Screen partition used for multi-GPU rendering — host-side mirror of FrameMap / slotToPixel / k_untile
(csrc/rtb_kernels.cuh, csrc/rtb_kernels.cu).

The frame is cut into 32x32-pixel blocks; block g belongs to rank g % nranks.  A rank's pixels live in "wavefront
slot" order: slot i -> local block k = i >> 10 (global block g = k * nranks + rank), 8x4 sub-tile s = (i >> 5) & 31,
lane = i & 31 -> pixel (bx*32 + (s & 3)*8 + (lane & 7), by*32 + (s >> 2)*4 + (lane >> 3)).  Every rank's tiled buffer
has slots_per_rank(w, h, n) words so a gather can use equal counts.  No reference counterpart: the reference is
single-GPU.
"""
from __future__ import annotations

import numpy as np

BLOCK = 32


def blocks(w: int, h: int):
    return (w + BLOCK - 1) // BLOCK, (h + BLOCK - 1) // BLOCK


def slots_per_rank(w: int, h: int, nranks: int) -> int:
    bx, by = blocks(w, h)
    return ((bx * by + nranks - 1) // nranks) * 1024


def local_blocks(w: int, h: int, rank: int, nranks: int) -> int:
    bx, by = blocks(w, h)
    total = bx * by
    return (total - rank + nranks - 1) // nranks if total > rank else 0


def slot_pixels(w: int, h: int, rank: int, nranks: int):
    """(x, y, valid) arrays over the rank's slots_per_rank slots; valid is False for slots outside the image or past the
    rank's last block."""
    n = slots_per_rank(w, h, nranks)
    i = np.arange(n, dtype=np.int64)
    k, s, lane = i >> 10, (i >> 5) & 31, i & 31
    g = k * nranks + rank
    bx, by = blocks(w, h)
    gx, gy = g % bx, g // bx
    x = gx * 32 + (s & 3) * 8 + (lane & 7)
    y = gy * 32 + (s >> 2) * 4 + (lane >> 3)
    valid = (g < bx * by) & (x < w) & (y < h)
    return x, y, valid


def untile(gathered: np.ndarray, w: int, h: int, nranks: int) -> np.ndarray:
    """gathered: (nranks, slots_per_rank) words -> (h, w) scan-line frame (what rtb_untile does on the GPU)."""
    out = np.zeros((h, w), gathered.dtype)
    for r in range(nranks):
        x, y, valid = slot_pixels(w, h, r, nranks)
        out[y[valid], x[valid]] = gathered[r][valid]
    return out
