"""Diagnostic: per-pass device time of every rank for the C4 tile-sharded frame (torchrun)."""
import importlib, os, sys, json
sys.path.insert(0, ".")
import torch, torch.distributed as dist
rt = importlib.import_module("rust-swift-raytracer_b200"); scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")
multi = importlib.import_module("rust-swift-raytracer_b200.multi")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = rt.load_world(scenes.default_world())
W, H = 3840, 2160
for gather in ("peer", "nccl"):
    r = multi.ShardedRenderer(rt, h, W, H, rank, world, tile_rows=16, gather=gather)
    for _ in range(2):
        r.render(1024, 8, 16)
    torch.cuda.synchronize(); dist.barrier()
    # one more frame, pass by pass, with events
    stream = torch.cuda.current_stream().cuda_stream
    accum = r.accum
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(17)]
    out_ptr = r.frame_ptr if gather == "peer" else r.local.data_ptr()
    evs[0].record()
    for p in range(16):
        last = p == 15
        o = rt.Options(64, 8, sample_begin=64 * p, resolve_spp=1024, tile_rows=16, shard_index=rank, shard_count=world,
                       accum_in=p > 0, accum_out=not last, no_resolve=not last, full_frame_out=(gather == "peer"))
        rt.render_device(h, o, W, H, out_ptr, accum.data_ptr(), stream)
        evs[p + 1].record()
    torch.cuda.synchronize()
    ms = [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(16)]
    box = [None] * world
    dist.all_gather_object(box, (rank, round(sum(ms), 1), ms))
    if rank == 0:
        for b in box:
            print(gather, b)
    dist.barrier()
    r.close()
dist.destroy_process_group()
