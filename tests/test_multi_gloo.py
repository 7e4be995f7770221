"""N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of the tile shard + gather plumbing
(rust-swift-raytracer_b200/multi.py) — the same code path bench.py uses over NCCL."""
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, W, H, tile_rows, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    multi = importlib.import_module("rust-swift-raytracer_b200.multi")
    rt = importlib.import_module("rust-swift-raytracer_b200")
    local = multi.alloc_compact(W, H, tile_rows, world, "cpu")
    # fill this rank's tiles with the global pixel index they must land on
    tiles = rt.shard_tiles(H, tile_rows, rank, world)
    assert len(tiles) == multi.tiles_of_rank(H, tile_rows, rank, world)
    assert rt.shard_pixel_count(W, H, tile_rows, rank, world) == multi.compact_pixels(W, H, tile_rows, rank, world)
    for j, (r0, r1) in enumerate(tiles):
        n = (r1 - r0) * W
        local[j * tile_rows * W: j * tile_rows * W + n] = torch.arange(r0 * W, r1 * W, dtype=torch.int32)
    frame = multi.gather_frame(local, W, H, tile_rows, rank, world)
    if rank == 0:
        q.put(frame.numpy().copy())
    else:
        assert frame is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H,tile_rows", [(2, 40, 36, 4), (2, 33, 50, 16), (3, 16, 23, 4), (2, 8, 3, 4)])
def test_tile_gather_reassembles_the_frame(world, W, H, tile_rows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, tile_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    frame = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(frame, np.arange(W * H, dtype=np.int32).reshape(H, W))


def test_single_rank_gather_is_identity():
    sys.path.insert(0, str(ROOT))
    multi = importlib.import_module("rust-swift-raytracer_b200.multi")
    local = multi.alloc_compact(10, 7, 4, 1, "cpu")
    local[:70] = torch.arange(70, dtype=torch.int32)
    f = multi.gather_frame(local, 10, 7, 4, 0, 1)
    assert f.shape == (7, 10) and int(f[6, 9]) == 69
