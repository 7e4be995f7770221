#!/usr/bin/env python
"""bench.py — throughput of the render hot path on B200 (BASELINE.json metric: Mrays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--fast-math]
    python bench.py --impl reference ...          # the reference's CPU algorithm (oracle port)

A "step" is one full frame of the workload (one pass of the per-pixel render loop,
common.rs:320-361, over every pixel and sample).  A "ray" is one ray segment = one World::hit
call (common.rs:268), counted by the kernel itself.

  N = 1 : BASELINE config 2 — default scene, 1920x1080, 64 spp, depth 8 — one persistent
          render-kernel launch per step.
  N > 1 : BASELINE config 4 — default scene, 3840x2160, 1,024 spp as 16 progressive passes of
          64 spp, the frame sharded by 16-row tiles across the ranks (tile t -> rank t % N), the
          finished RGBA8 tiles gathered to rank 0 over NCCL/NVLink (the only collective).
          Strong scaling: the frame is fixed, N GPUs split it.

`value`  : rays/s with everything resident in HBM, timed with CUDA events on the launching
           stream, L2 flushed (256 MiB write) between steps outside the events, max over ranks.
`e2e`    : the same frame through the reference-facing C-ABI call with a HOST framebuffer
           (render_with_options -> D2H of the finished frame inside the timed region).
`roofline`: FP32 CUDA-core roofline (no tensor-core work exists on this path): algorithmic
           flops (DESIGN.md / SURVEY.md §8d) / kernel time / FFMA peak measured in this run.
The default kernel is the bit-exact one (IEEE arithmetic in the reference's association order,
no FMA contraction); --fast-math selects the relaxed kernel, which is not bit-exact.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # key: (description, scene, W, H, spp, depth, passes)
    "c1": ("C1: default scene (world.txt, 8 spheres), 400x224, 50 spp, depth 8", "default", 400, 224, 50, 8, 1),
    "c2": ("C2: default scene (world.txt, 8 spheres), 1920x1080, 64 spp, depth 8", "default", 1920, 1080, 64, 8, 1),
    "c3": ("C3: synthetic 1,000 spheres, 1920x1080, 256 spp, depth 8", "c3", 1920, 1080, 256, 8, 1),
    "c4": ("C4: default scene, 3840x2160, 1,024 spp (16 progressive passes x 64), depth 8, 16-row tiles",
           "default", 3840, 2160, 1024, 8, 16),
    "c5": ("C5: synthetic 8,000 spheres + 2,000 triangles, 1280x720, 16 spp, depth 16", "c5", 1280, 720, 16, 16, 1),
}


# dram__bytes_read.sum + dram__bytes_write.sum of one render-kernel launch, from the committed
# `ncu --set full` captures (profiles/r01_bench.md); None where no capture of that exact launch exists.
NCU_TRAFFIC_BYTES = {("c2", False): 36352}      # 36 KB read, 0 B written: the 8.3 MB frame stays in the 126 MB L2
# Counters of the same captures, quoted (not measured in this run) so that the line explains its own
# `frac`: on the 8-sphere scene the counted-flops model covers less than half of the instructions a
# segment needs (IEEE divide/sqrt sequences, integer RNG, predicates), and the kernel is issue-bound.
NCU_CONTEXT = {
    ("c2", False): {"issue_slots_busy": 0.897, "ipc": 3.59, "fp32_pipe_active": 0.443, "active_lanes_per_warp": 22.4,
                    "source": "profiles/r01_c2_exact_v3_full_size_light.txt, r01_c2_exact_v4_spp16_details.txt"},
    ("c3", False): {"issue_slots_busy": 0.728, "ipc": 2.91, "fp32_pipe_active": 0.500, "active_lanes_per_warp": 30.4,
                    "source": "profiles/r01_c3_exact_v4_spp8_details.txt (64-register build)"},
}


def scene_text(scenes, key):
    return {"default": scenes.default_world, "c3": scenes.c3_world, "c5": scenes.c5_world}[key]()


def algorithmic_flops(rays, samples, pixels, n_sph, n_tri):
    """N_seg*(17*S + 12*T + 60) + N_s*34 + W*H*11, FMA counted as 2."""
    return rays * (17 * n_sph + 12 * n_tri + 60) + samples * 34 + pixels * 11


class ClockSampler:
    """Samples SM clock and clock-event reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:       # noqa: BLE001
            self.nv, self.err = None, repr(e)
        self.period = period

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake",
                 nv.nvmlClocksEventReasonApplicationsClocksSetting: "applications_clocks_setting"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:        # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------- CPU legs

def cpu_baseline_serial(scenes, wl, budget_s=15.0):
    """The reference's algorithm (oracle port, serial RNG stream, 1 thread) on a bounded sample of
    the workload: the full frame at reduced spp (cost is exactly linear in spp)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_binding as ob
    desc, key, W, H, spp, depth, _ = wl
    cam, world = ob.parse_input(scene_text(scenes, key))
    # size the sample from a short probe
    probe_h = max(2, H // 16)
    t = time.perf_counter()
    _, r0, _ = ob.ray_trace(world, cam, W, probe_h, 1, depth, rng_mode=ob.RNG_SERIAL, threads=1)
    per_spp = (time.perf_counter() - t) * (H / probe_h)
    s = int(max(1, min(spp, budget_s / max(per_spp, 1e-6))))
    t = time.perf_counter()
    _, rays, _ = ob.ray_trace(world, cam, W, H, s, depth, rng_mode=ob.RNG_SERIAL, threads=1)
    dt = time.perf_counter() - t
    ncores = os.cpu_count() or 1
    t = time.perf_counter()
    _, rays_mt, _ = ob.ray_trace(world, cam, W, H, s, depth, rng_mode=ob.RNG_PER_SAMPLE, threads=ncores)
    dt_mt = time.perf_counter() - t
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
            "sample": f"{W}x{H} full frame at {s} of {spp} spp, depth {depth}; oracle/rt_oracle.c serial-RNG mode "
                      f"(the reference's one xorshift32 stream, common.rs:321), {dt:.1f} s; extrapolated frame time "
                      f"{dt * spp / s:.0f} s",
            "ms_per_frame_extrapolated": dt * spp / s * 1e3,
            "all_cores": {"value": rays_mt / dt_mt / 1e6, "cores": ncores,
                          "note": "oracle per-sample-RNG mode, OpenMP over rows (not something the reference can do)"}}


def run_reference(args, scenes):
    """--impl reference: the reference's CPU implementation of the path.  The Rust crate cannot be
    built in this image (no rustc/cargo), so this is the oracle port in the reference's own
    serial-RNG mode on 1 thread — the reference's render loop is single-threaded by construction
    (one &mut Random shared by every pixel, common.rs:321-340)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_binding as ob
    wl = WORKLOADS[args.workload]
    desc, key, W, H, spp, depth, _ = wl
    cam, world = ob.parse_input(scene_text(scenes, key))
    # one step = the full frame at `s` spp, sized so that (K + W) steps take <= ~60 s
    probe_h = max(2, H // 16)
    t = time.perf_counter()
    ob.ray_trace(world, cam, W, probe_h, 1, depth, rng_mode=ob.RNG_SERIAL, threads=1)
    per_spp = (time.perf_counter() - t) * (H / probe_h)
    s = int(max(1, min(spp, 60.0 / max(per_spp * (args.steps + args.warmup), 1e-6))))
    for _ in range(args.warmup):
        ob.ray_trace(world, cam, W, H, s, depth, rng_mode=ob.RNG_SERIAL, threads=1)
    rays_total, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        _, rays, _ = ob.ray_trace(world, cam, W, H, s, depth, rng_mode=ob.RNG_SERIAL, threads=1)
        rays_total += rays
    dt = time.perf_counter() - t0
    value = rays_total / dt / 1e6
    ncores = os.cpu_count() or 1
    t = time.perf_counter()
    _, rays_mt, _ = ob.ray_trace(world, cam, W, H, s, depth, rng_mode=ob.RNG_PER_SAMPLE, threads=ncores)
    dt_mt = time.perf_counter() - t
    sample = (f"each step = {W}x{H} full frame at {s} of {spp} spp, depth {depth}, oracle/rt_oracle.c in the "
              f"reference's serial-RNG mode, 1 thread (the reference is single-threaded)")
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "note": "rays/s is independent of spp; frame time scales linearly in spp"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 1, "kind": "port", "sample": sample,
                             "all_cores": {"value": rays_mt / dt_mt / 1e6, "cores": ncores,
                                           "note": "oracle per-sample-RNG mode with OpenMP over rows"}},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--fast-math", action="store_true", help="relaxed-arithmetic kernel (not bit-exact)")
    ap.add_argument("--spp", type=int, default=None, help="override the workload's spp (profiling only: "
                    "the line then says so in config.workload)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = render kernels store their tiles into rank 0's frame over NVLink "
                         "(CUDA IPC mapping); 'nccl' = compact buffers + one dist.gather")
    ap.add_argument("--devices", type=int, default=0,
                    help="N = 1 launch only: e2e through render_with_options(n_devices=D), ONE process driving D GPUs "
                         "(the shape the C-ABI callers have); reported under e2e_one_process")
    ap.add_argument("--cull", action="store_true",
                    help="opt-in acceleration mode (RT_OPT_GROUP_CULL): same frame, fewer sphere tests — a separately "
                         "reported mode, NOT the brute-force path the BASELINE metric is defined on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = "c2" if max(args.gpus, world_size) == 1 else "c4"
    heavy = args.workload in ("c3", "c4")
    if args.steps is None:
        args.steps = 5 if heavy else 50
    if args.warmup is None:
        args.warmup = 3 if heavy else 5

    if args.impl != "reference":
        args.warmup = max(args.warmup, 3)            # timing rule: at least 3 untimed warm-up steps
    args.steps = max(args.steps, 1)

    scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")
    if args.impl == "reference":
        return run_reference(args, scenes)

    import torch
    import torch.distributed as dist
    build = importlib.import_module("rust-swift-raytracer_b200.build")
    rt = importlib.import_module("rust-swift-raytracer_b200")
    if not rt.LIB_PATH.exists():
        build.build()
    multi = importlib.import_module("rust-swift-raytracer_b200.multi")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or rt.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world_size
    if args.gpus != n_gpus and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world_size}; using {n_gpus}", file=sys.stderr)

    wl = WORKLOADS[args.workload]
    desc, key, W, H, spp, depth, passes = wl
    if args.spp:
        spp, passes = args.spp, 1
        desc += f" [REDUCED to {spp} spp for profiling — not a bench line]"
    handle = rt.load_world(scene_text(scenes, key))
    S, T = handle.n_spheres, handle.n_triangles
    dev = torch.device("cuda", local_rank)
    renderer = multi.ShardedRenderer(rt, handle, W, H, rank, n_gpus, tile_rows=16, device=dev, gather=args.gather)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(x, op):
        if n_gpus == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    fast = args.fast_math
    # FP32 peak of this box, measured before the timed region (FFMA-chain microbenchmark)
    fp32_peak = rt.measure_fp32_peak(local_rank) if rank == 0 else 0.0

    # ---- warm-up (first step also counts the rays of one frame; the frame is deterministic) ----
    _, rays_local = renderer.render(spp, depth, passes, fast_math=fast, count_rays=True, group_cull=args.cull)
    rays_frame = allreduce(float(rays_local), dist.ReduceOp.SUM if n_gpus > 1 else None)
    for _ in range(max(args.warmup - 1, 0)):
        flush.zero_()
        renderer.render(spp, depth, passes, fast_math=fast, group_cull=args.cull)
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream around every step ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall = time.perf_counter()
        for a, b in ev:
            flush.zero_()                       # L2 flush, outside the events
            a.record()
            renderer.render(spp, depth, passes, fast_math=fast, group_cull=args.cull)
            b.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    dev_ms = allreduce(dev_ms, dist.ReduceOp.MAX if n_gpus > 1 else None)
    ms_per_step = dev_ms / args.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host framebuffer through the public API, D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        if n_gpus == 1:
            fb = rt.Framebuffer(W, H, pinned=True)
            opts = rt.Options(spp, depth, fast_math=fast, group_cull=args.cull)

            def e2e_step():
                if passes == 1:
                    rt.render_with_options(fb, handle, opts)     # the reference-facing C-ABI call
                else:
                    renderer.render(spp, depth, passes, fast_math=fast, to_host=True, group_cull=args.cull)
        else:
            def e2e_step():
                renderer.render(spp, depth, passes, fast_math=fast, to_host=True, group_cull=args.cull)
        e2e_step()
        barrier()
        tot = 0.0
        for _ in range(args.steps):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            e2e_step()
            torch.cuda.synchronize()
            tot += allreduce(time.perf_counter() - t0, dist.ReduceOp.MAX if n_gpus > 1 else None)
        e2e = {"value": rays_frame / (tot / args.steps) / 1e6, "unit": "Mrays/s",
               "ms_per_frame": tot / args.steps * 1e3,
               "h2d_bytes_per_step": 232 * passes,      # camera + frame parameters travel as kernel arguments
               # (sizeof RtFrameParams + RtSceneView = 160 + 72);
               # the scene blob is uploaded once by load_world (the reference's API has the same split)
               "d2h_bytes_per_step": W * H * 4,
               "api": "render_with_options (C ABI, pinned host framebuffer: the kernel stores the finished pixels "
                      "straight into it over PCIe, no separate D2H copy)" if (n_gpus == 1 and passes == 1)
                      else f"multi.ShardedRenderer.render(to_host=True): tile shards -> {renderer.gather} gather -> D2H on rank 0"}

    one_process = None
    if args.devices > 1 and n_gpus == 1 and passes == 1:
        fbm = rt.Framebuffer(W, H, pinned=True)
        om = rt.Options(spp, depth, fast_math=fast, n_devices=args.devices, group_cull=args.cull)
        stm = rt.RenderStats()
        rt.render_with_options(fbm, handle, om, stm)
        tot = 0.0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            rt.render_with_options(fbm, handle, om)
            tot += time.perf_counter() - t0
        one_process = {"devices": int(stm.devices), "peer_gather": int(stm.peer_gather),
                       "value": rays_frame / (tot / args.steps) / 1e6, "unit": "Mrays/s",
                       "ms_per_frame": tot / args.steps * 1e3, "d2h_bytes_per_step": W * H * 4,
                       "api": "render_with_options(n_devices=D): one process, tiles stored into device 0's frame "
                              "over NVLink peer mappings, one D2H"}
    renderer.close()
    if rank != 0:
        if n_gpus > 1:
            dist.destroy_process_group()
        return 0

    samples = W * H * spp
    flops = algorithmic_flops(rays_frame, samples, W * H, S, T)
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if n_gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "spheres": S, "triangles": T, "rays_per_step": int(rays_frame),
                   "samples_per_step": samples, "kernel": "fast-math" if fast else "exact (bit-identical to the oracle)",
                   "sphere_walk": ("group-cull (opt-in acceleration: the roofline fraction below still counts the brute-force "
                                   "algorithmic flops, so it is a speed-up measure, not a pipe utilisation)") if args.cull
                                  else "brute force over the list (filtered for >= 64 spheres)",
                   "parallelism": (f"row-tile shards x{n_gpus}, {renderer.gather} gather to rank 0 "
                                   + ("(tiles stored by the render kernels into rank 0's frame over NVLink, CUDA IPC)"
                                      if renderer.gather == "peer" else "(compact buffers + dist.gather)"))
                   if n_gpus > 1 else "1 GPU",
                   "l2": "flushed between steps (256 MiB write, outside the CUDA events)",
                   "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
                   },
        "e2e": e2e,
        "gpu_launches": args.steps * passes,
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak * n_gpus, "unit": "TFLOP/s",
                     "frac": achieved / (fp32_peak * n_gpus) if fp32_peak else None,
                     "traffic": NCU_TRAFFIC_BYTES.get((args.workload, fast)) if (n_gpus == 1 and not args.spp) else None,
                     "peak_source": "FFMA-chain microbenchmark run on this GPU before the timed region "
                                    "(MEASURED_PEAKS.json has no FP32 CUDA-core figure; nominal 2*128*148*1.965 GHz = 74.4)",
                     "flops_per_step": flops,
                     "flops_model": "rays*(17*S + 12*T + 60) + samples*34 + pixels*11, FMA = 2 (SURVEY.md 8d)",
                     "hbm_bytes_per_step": W * H * 4 + (passes - 1) * 2 * W * H * 16,
                     "ncu_context": NCU_CONTEXT.get((args.workload, fast)) if n_gpus == 1 else None},
        "clocks": clocks.summary(),
    }
    if one_process:
        line["e2e_one_process"] = one_process
    if not args.no_cpu_baseline and n_gpus == 1:
        line["cpu_baseline"] = cpu_baseline_serial(scenes, wl)
    print(json.dumps(line), flush=True)
    if n_gpus > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
