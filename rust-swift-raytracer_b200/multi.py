"""Multi-GPU plumbing of the render path: one process per GPU, the frame sharded by row
tiles, one collective (the gather of finished RGBA8 tiles to rank 0).

Decomposition: the frame is cut into tiles of `tile_rows` image rows; tile t
belongs to rank t % world (static interleave: sky-heavy and geometry-heavy tiles alternate
across ranks).  Every tile is one contiguous byte range of the row-major frame (image.rs:27),
so a rank's output is its tiles packed back to back ("compact" buffer, RT_FLAG_COMPACT_OUT)
and the gather is a pure data movement: no pixel is touched by more than one rank and no
float accumulator ever crosses NVLink.

torch.distributed is plumbing here (NCCL on the GPUs, gloo in the CPU tests); nothing in this
module computes pixels.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def tiles_total(height: int, tile_rows: int) -> int:
    return (height + tile_rows - 1) // tile_rows


def tiles_of_rank(height: int, tile_rows: int, rank: int, world: int) -> int:
    t = tiles_total(height, tile_rows)
    return 0 if rank >= t else (t - rank + world - 1) // world


def compact_pixels(width: int, height: int, tile_rows: int, rank: int, world: int) -> int:
    """Pixels in rank's compact buffer (every tile padded to tile_rows rows)."""
    return tiles_of_rank(height, tile_rows, rank, world) * tile_rows * width


def padded_tiles_per_rank(height: int, tile_rows: int, world: int) -> int:
    return (tiles_total(height, tile_rows) + world - 1) // world


def alloc_compact(width: int, height: int, tile_rows: int, world: int, device) -> torch.Tensor:
    """A rank's output buffer, padded to the largest shard so the gather is uniform."""
    j = padded_tiles_per_rank(height, tile_rows, world)
    return torch.zeros(j * tile_rows * width, dtype=torch.int32, device=device)


def gather_frame(local: torch.Tensor, width: int, height: int, tile_rows: int, rank: int, world: int,
                 staging: Optional[torch.Tensor] = None, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather every rank's compact RGBA8 buffer (int32 per pixel) to `dst` and interleave the
    tiles back into frame order.  Returns the [height, width] int32 frame on dst, None elsewhere.

    `local` must come from alloc_compact (uniform size).  `staging` ([world, J*tile_px] int32 on
    dst) can be passed to avoid re-allocation between frames.
    """
    tile_px = tile_rows * width
    j = padded_tiles_per_rank(height, tile_rows, world)
    assert local.numel() == j * tile_px, (local.numel(), j, tile_px)
    if world == 1:
        return local[: width * height].view(height, width)
    if rank == dst:
        if staging is None:
            staging = torch.empty((world, j * tile_px), dtype=local.dtype, device=local.device)
        dist.gather(local, gather_list=list(staging.unbind(0)), dst=dst, group=group)
        # staging[r, jj] is tile jj*world + r  ->  frame order is [jj][r]
        frame = staging.view(world, j, tile_px).permute(1, 0, 2).reshape(-1)[: width * height]
        return frame.view(height, width)
    dist.gather(local, gather_list=None, dst=dst, group=group)
    return None
