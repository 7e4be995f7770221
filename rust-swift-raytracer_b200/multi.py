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

    Every rank renders its row tiles of the frame (rt_shard_tile) — progressive passes fused into one
    persistent launch, the float4 sums never leaving the GPU — and the finished RGBA8 tiles end up on
    rank 0.  Two gathers:

      gather="peer" (default for world > 1): rank 0 owns the frame (rt_device_alloc) and exports
          it through CUDA IPC; the other ranks map it and their render kernels STORE their tiles
          straight into rank 0's memory over NVLink (RT_OPT_FULL_FRAME_OUT) — the gather is
          fused into the pack step of the kernel.  A one-element all-reduce enqueued after the
          kernels orders rank 0's read after everybody's stores; it carries no frame data.
          With steal=True every rank also exports its shard block (work counter + sums); a GPU whose
          own queue is empty takes slabs of the other GPUs' queues over NVLink (static deal + work
          stealing of the tail, rt_types.h RtQueue).
      gather="nccl": every rank packs its tiles into a compact buffer; one dist.gather to rank 0
          and a permuting view reassemble the frame (also the gloo/CPU-testable path).

    All kernels, the collective and the optional D2H are enqueued on torch's current CUDA
    stream, so CUDA events recorded on that stream bracket the whole step.

    Frame lifetime (peer gather): the frame returned by render() — the pinned host tensor or the raw
    device address — is valid until the NEXT render() call on any rank; every render() therefore
    starts with a token all-reduce, which rank 0 joins only after its previous D2H, so no rank can
    store pixels of frame k+1 into a frame rank 0 is still copying out.
    """

    def __init__(self, rt, handle, width: int, height: int, rank: int = 0, world: int = 1,
                 tile_rows: int = 16, device=None, gather: Optional[str] = None, steal: bool = True,
                 row_gather: bool = True):
        self.rt, self.handle = rt, handle
        self.width, self.height, self.rank, self.world, self.tile_rows = width, height, rank, world, tile_rows
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.gather = gather or ("peer" if world > 1 else "nccl")
        assert self.gather in ("peer", "nccl")
        self.host_frame = (torch.empty((height, width), dtype=torch.int32).pin_memory() if rank == 0 else None)
        self.frame_ptr = 0        # peer mode: rank 0's frame (own allocation on rank 0, IPC mapping elsewhere)
        self._owns_frame = False
        self.peer_error = None
        self.block_ptr = 0        # this rank's shard block (work stealing)
        self._peer_blocks = {}    # rank -> mapped address of its shard block
        self.queues = None        # [(block address, shard index)] own first, then (rank+1), (rank+2), ...
        self.steal = False        # cross-GPU work stealing is on
        self.row_gather = bool(row_gather)   # ranks != 0 render into a local frame; a copy kernel moves it over as 16-byte vectors
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
                if steal and world <= 8 and not os.environ.get("RT_DISABLE_STEAL"):
                    self._exchange_blocks()
        if not (self.gather == "peer" and world > 1):
            self.gather = "nccl"
            self.local = alloc_compact(width, height, tile_rows, world, self.device)
            j = padded_tiles_per_rank(height, tile_rows, world)
            self.staging = (torch.empty((world, j * tile_rows * width), dtype=torch.int32, device=self.device)
                            if (world > 1 and rank == 0) else None)

    def _exchange_blocks(self):
        """Every rank allocates its shard block, marks its queue empty, and maps everybody else's."""
        rt = self.rt
        ok, mine = 1, None
        try:
            self.block_ptr = rt.device_alloc(rt.shard_block_bytes(self.width, self.height))
            rt.shard_block_init(self.block_ptr)
            mine = rt.ipc_export(self.block_ptr)
        except rt.RenderError as e:
            ok, self.peer_error = 0, str(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        if ok:
            try:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    if h is None:
                        raise rt.RenderError(f"rank {r} could not export its shard block")
                    self._peer_blocks[r] = rt.ipc_open(h)
            except rt.RenderError as e:
                ok, self.peer_error = 0, str(e)
        t = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)           # also: every block is initialised before anyone may raid it
        if int(t.item()) == 0:
            self._release_blocks()
            return
        order = [(self.rank + i) % self.world for i in range(self.world)]
        self.queues = [((self.block_ptr if r == self.rank else self._peer_blocks[r]), r) for r in order]
        self.steal = True

    def _release_blocks(self):
        for p in self._peer_blocks.values():
            self.rt.ipc_close(p)
        self._peer_blocks = {}
        self.queues, self.steal = None, False

    def close(self):
        if self.frame_ptr or self.block_ptr:
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier()                # nobody may still be storing into the frame or raiding a queue
            self._release_blocks()
            if self.frame_ptr and not self._owns_frame:
                self.rt.ipc_close(self.frame_ptr)
            if self.world > 1:
                dist.barrier()                # all mappings are closed before their owners free the memory
            if self.frame_ptr and self._owns_frame:
                self.rt.device_free(self.frame_ptr)
            if self.block_ptr:
                self.rt.device_free(self.block_ptr)
            self.frame_ptr = self.block_ptr = 0

    def render(self, spp: int, depth: int, passes: int = 1, seed: Optional[int] = None, fast_math: bool = False,
               fixed_jitter: bool = False, to_host: bool = False, count_rays: bool = False, group_cull: bool = False,
               stats=None):
        """Returns (frame, rays).  frame: on rank 0 the [H, W] int32 RGBA8 frame — the pinned host
        tensor when to_host, else a device tensor (nccl gather) or the raw device address of the
        frame (peer gather); None on the other ranks.  rays: this rank's ray-segment count when
        count_rays (that mode synchronises the stream), else 0."""
        rt = self.rt
        assert passes >= 1 and spp % passes == 0, "spp must divide evenly into passes"
        peer = self.gather == "peer"
        tstream = torch.cuda.current_stream(self.device)
        stream = tstream.cuda_stream
        if peer:
            dist.all_reduce(self.token)       # frame-reuse fence (see the class docstring)
        out_ptr = self.frame_ptr if peer else self.local.data_ptr()
        o = rt.Options(spp, depth, passes=passes, resolve_spp=spp, fast_math=fast_math, fixed_jitter=fixed_jitter,
                       tile_rows=self.tile_rows, shard_index=self.rank, shard_count=self.world,
                       full_frame_out=peer, group_cull=group_cull, peer_queues=self.queues if self.steal else None,
                       row_gather=peer and self.rank != 0 and self.row_gather)
        if seed is not None:
            o.seed = seed
        st = stats if stats is not None else (rt.RenderStats() if count_rays else None)
        rt.render_device(self.handle, o, self.width, self.height, out_ptr, 0, stream, st)
        rays = st.rays if st is not None else 0
        if peer:
            # stream-ordered completion fence: rank 0's all-reduce kernel cannot finish before every
            # rank has launched its own, i.e. before every rank's render kernel has completed
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
