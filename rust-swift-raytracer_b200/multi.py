"""Multi-GPU plumbing of the render path: one process per GPU, the frame sharded by row
tiles, one collective (the gather of finished RGBA8 tiles to rank 0).

Decomposition: the frame is cut into tiles of `tile_rows` image rows, dealt to the
ranks in stripes of `world` consecutive tiles, alternately forwards and backwards (rt_shard_tile,
rt_types.h) so that the frame's vertical cost gradient (sky above, ground below) cancels between
the ranks.  Every tile is one contiguous byte range of the row-major frame (image.rs:27), so a
rank's output is either its tiles packed back to back ("compact" buffer, RT_FLAG_COMPACT_OUT) or
stores into the full frame at the tiles' offsets; no pixel is touched by more than one rank and no
float accumulator ever crosses NVLink.

torch.distributed is plumbing here (NCCL on the GPUs, gloo in the CPU tests); nothing in this
module computes pixels.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def tiles_total(height: int, tile_rows: int) -> int:
    return (height + tile_rows - 1) // tile_rows


def tiles_of_rank(height: int, tile_rows: int, rank: int, world: int) -> int:
    """Tiles of `rank` under the boustrophedon dealing of rt_shard_tile (rt_types.h)."""
    t = tiles_total(height, tile_rows)
    full, rem = divmod(t, world)
    pos = (world - 1 - rank) if (full & 1) else rank
    return full + (1 if pos < rem else 0)


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
        # staging[r, jj] is tile jj*world + (r if jj is even else world-1-r) (rt_shard_tile): frame order is
        # [jj][r], with the rank axis reversed in the odd stripes
        tiles = staging.view(world, j, tile_px).permute(1, 0, 2).clone()
        tiles[1::2] = tiles[1::2].flip(1)
        frame = tiles.reshape(-1)[: width * height]
        return frame.view(height, width)
    dist.gather(local, gather_list=None, dst=dst, group=group)
    return None


class ShardedRenderer:
    """One rank's share of a tile-sharded frame, device-resident.

    Every rank renders tiles rank, rank+world, ... of the frame — optionally as several
    progressive passes through a float4 accumulator that never leaves the GPU — and the finished
    RGBA8 tiles end up on rank 0.  Two gathers:

      gather="peer" (default for world > 1): rank 0 owns the frame (rt_device_alloc) and exports
          it through CUDA IPC; the other ranks map it and their render kernels STORE their tiles
          straight into rank 0's memory over NVLink (RT_OPT_FULL_FRAME_OUT) — the gather is
          fused into the pack step of the kernel.  A one-element all-reduce enqueued after the
          kernels orders rank 0's read after everybody's stores; it carries no frame data.
      gather="nccl": every rank packs its tiles into a compact buffer; one dist.gather to rank 0
          and a permuting view reassemble the frame (also the gloo/CPU-testable path).

    All kernels, the collective and the optional D2H are enqueued on torch's current CUDA
    stream, so CUDA events recorded on that stream bracket the whole step.
    """

    def __init__(self, rt, handle, width: int, height: int, rank: int = 0, world: int = 1,
                 tile_rows: int = 16, device=None, gather: Optional[str] = None):
        self.rt, self.handle = rt, handle
        self.width, self.height, self.rank, self.world, self.tile_rows = width, height, rank, world, tile_rows
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.gather = gather or ("peer" if world > 1 else "nccl")
        assert self.gather in ("peer", "nccl")
        self.accum: Optional[torch.Tensor] = None
        self.host_frame = (torch.empty((height, width), dtype=torch.int32).pin_memory() if rank == 0 else None)
        self.frame_ptr = 0        # peer mode: rank 0's frame (own allocation on rank 0, IPC mapping elsewhere)
        self._owns_frame = False
        self.peer_error = None
        if self.gather == "peer" and world > 1:
            # Map rank 0's frame into every rank.  CUDA IPC can be unavailable (containers without a
            # shared IPC namespace, no peer access): all ranks then agree to use the NCCL gather — a
            # slower data path with the same result, never a different computation.
            nbytes = width * height * 4
            box, ok = [None], 1
            try:
                if rank == 0:
                    self.frame_ptr = rt.device_alloc(nbytes)
                    self._owns_frame = True
                    box[0] = rt.ipc_export(self.frame_ptr)
            except rt.RenderError as e:
                ok, self.peer_error = 0, str(e)
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                try:
                    if box[0] is None:
                        raise rt.RenderError("rank 0 could not export its frame")
                    if os.environ.get("RT_DISABLE_IPC"):      # test hook for the fallback below
                        raise rt.RenderError("CUDA IPC disabled by RT_DISABLE_IPC")
                    self.frame_ptr = rt.ipc_open(box[0])
                except rt.RenderError as e:
                    ok, self.peer_error = 0, str(e)
            self.token = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(self.token, op=dist.ReduceOp.MIN)
            if int(self.token.item()) == 0:
                if self.frame_ptr:
                    (rt.device_free if self._owns_frame else rt.ipc_close)(self.frame_ptr)
                self.frame_ptr, self._owns_frame = 0, False
                self.gather = "nccl"
            else:
                self.token.zero_()
                self.local = None
                self.staging = None
        if not (self.gather == "peer" and world > 1):
            self.gather = "nccl"
            self.local = alloc_compact(width, height, tile_rows, world, self.device)
            j = padded_tiles_per_rank(height, tile_rows, world)
            self.staging = (torch.empty((world, j * tile_rows * width), dtype=torch.int32, device=self.device)
                            if (world > 1 and rank == 0) else None)

    def close(self):
        if self.frame_ptr:
            torch.cuda.synchronize(self.device)
            if self._owns_frame:
                if self.world > 1:
                    dist.barrier()            # nobody may still be storing into the frame
                self.rt.device_free(self.frame_ptr)
            else:
                self.rt.ipc_close(self.frame_ptr)
                dist.barrier()
            self.frame_ptr = 0

    def render(self, spp: int, depth: int, passes: int = 1, seed: Optional[int] = None, fast_math: bool = False,
               fixed_jitter: bool = False, to_host: bool = False, count_rays: bool = False, group_cull: bool = False):
        """Returns (frame, rays).  frame: on rank 0 the [H, W] int32 RGBA8 frame — the pinned host
        tensor when to_host, else a device tensor (nccl gather) or the raw device address of the
        frame (peer gather); None on the other ranks.  rays: this rank's ray-segment count when
        count_rays (that mode synchronises after every pass), else 0."""
        rt = self.rt
        assert passes >= 1 and spp % passes == 0, "spp must divide evenly into passes"
        peer = self.gather == "peer"
        if passes > 1 and self.accum is None:
            n = self.width * self.height if peer else self.local.numel()
            self.accum = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        tstream = torch.cuda.current_stream(self.device)
        stream = tstream.cuda_stream
        out_ptr = self.frame_ptr if peer else self.local.data_ptr()
        per, rays = spp // passes, 0
        for p in range(passes):
            last = p == passes - 1
            o = rt.Options(per, depth, sample_begin=p * per, resolve_spp=spp, fast_math=fast_math,
                           fixed_jitter=fixed_jitter, tile_rows=self.tile_rows, shard_index=self.rank,
                           shard_count=self.world, accum_in=p > 0, accum_out=not last, no_resolve=not last,
                           full_frame_out=peer, group_cull=group_cull)
            if seed is not None:
                o.seed = seed
            st = rt.RenderStats() if count_rays else None
            rt.render_device(self.handle, o, self.width, self.height, out_ptr,
                             self.accum.data_ptr() if self.accum is not None else 0, stream, st)
            if st is not None:
                rays += st.rays
        if peer:
            # stream-ordered completion fence: rank 0's all-reduce kernel cannot finish before every
            # rank has launched its own, i.e. before every rank's render kernels have completed
            dist.all_reduce(self.token)
            frame = self.frame_ptr if self.rank == 0 else None
            if to_host and self.rank == 0:
                rt.copy_to_host(self.host_frame.data_ptr(), self.frame_ptr, self.width * self.height * 4, stream)
                tstream.synchronize()
                frame = self.host_frame
            return frame, rays
        frame = gather_frame(self.local, self.width, self.height, self.tile_rows, self.rank, self.world,
                             staging=self.staging)
        if to_host and frame is not None:
            self.host_frame.copy_(frame, non_blocking=True)
            tstream.synchronize()
            frame = self.host_frame
        return frame, rays
