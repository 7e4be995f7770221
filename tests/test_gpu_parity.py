"""Parity tests proper (B200 only, `-m gpu`): the CUDA render path, called through the C ABI
of libraytracer.so, against the CPU oracle on the same seeded inputs, against the committed
golden frames, and — at BASELINE.json's full sizes — through size-independent properties.

Bars (north_star):
  * exact kernel (default): bit-exact RGBA8 and identical ray-segment counts vs the oracle
    in per-sample RNG mode — stricter than the stated "max abs <= 1/255 after pack";
  * fast-math kernel and the reference's serial-RNG realisation: statistical agreement, RMSE
    bound calibrated from the noise floor of two independent oracle renders (stated below).
"""
import ctypes as C
import hashlib
from pathlib import Path

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = np.load(ROOT / "tests" / "golden" / "frames.npz")


def _render(rt, handle, W, H, spp, depth, pinned=False, **kw):
    fb = rt.Framebuffer(W, H, pinned=pinned)
    st = rt.RenderStats()
    rt.render_with_options(fb, handle, rt.Options(spp, depth, **kw), st)
    return fb.pixels.copy(), st


def _rmse(a, b):
    d = a[:, :, :3].astype(np.float64) - b[:, :, :3].astype(np.float64)
    return float(np.sqrt((d * d).mean()))


# ------------------------------------------------------------------ exact kernel == oracle

@pytest.mark.parametrize("case", cases.SMALL_CASES, ids=[c[0] for c in cases.SMALL_CASES])
def test_exact_kernel_equals_oracle_and_golden(gpu_rt, ob, scenes, case):
    rt = gpu_rt
    name, key, camera, W, H, spp, depth, fixed = case
    h = cases.product_scene(rt, scenes, key, camera)
    cam, world = cases.oracle_scene(ob, scenes, key, camera)
    assert np.array_equal(h.camera_floats(), cam.floats())
    got, st = _render(rt, h, W, H, spp, depth, fixed_jitter=fixed)
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth, fixed_jitter=fixed)
    assert st.launches == 1 and st.rays == rays            # same number of World::hit calls
    diff = np.abs(got.astype(int) - want.astype(int))
    assert diff.max() == 0, f"{name}: max abs {diff.max()} (tolerance: bit-exact; north_star allows 1)"
    assert np.array_equal(got, GOLDEN[name]) and rays == int(GOLDEN[name + "__rays"][0])


@pytest.mark.parametrize("case", cases.SMALL_CASES, ids=[c[0] for c in cases.SMALL_CASES])
def test_sample_item_scheduling_is_bit_identical(gpu_rt, ob, scenes, case):
    """RT_OPT_SAMPLE_ITEMS: lanes trace single samples, a second kernel adds them in sample order.
    Pure scheduling — pixels and ray counts must not change by a bit; nor with whole-pixel items
    forced (RT_OPT_PIXEL_ITEMS)."""
    rt = gpu_rt
    name, key, camera, W, H, spp, depth, fixed = case
    h = cases.product_scene(rt, scenes, key, camera)
    got, st = _render(rt, h, W, H, spp, depth, fixed_jitter=fixed, sample_items=True)
    assert st.sample_items == 1 and st.launches == 2
    assert np.array_equal(got, GOLDEN[name]) and st.rays == int(GOLDEN[name + "__rays"][0])
    got, st = _render(rt, h, W, H, spp, depth, fixed_jitter=fixed, sample_items=False)
    assert st.sample_items == 0 and st.launches == 1
    assert np.array_equal(got, GOLDEN[name]) and st.rays == int(GOLDEN[name + "__rays"][0])


def test_sample_items_with_shards_and_progressive_passes(gpu_rt, ob, scenes):
    import torch
    rt = gpu_rt
    W, H = 150, 70
    h = rt.load_world(scenes.example_world())
    cam, world = ob.parse_input(scenes.example_world())
    want, rays, want_acc = ob.ray_trace(world, cam, W, H, 12, 8, want_accum=True)
    # shards into the caller's frame
    fb = rt.Framebuffer(W, H)
    for i in range(3):
        rt.render_with_options(fb, h, rt.Options(12, 8, shard_index=i, shard_count=3, tile_rows=8, sample_items=True))
    assert np.array_equal(fb.pixels, want)
    # 3 progressive passes of 4 spp through the float4 accumulator
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    for p in range(3):
        o = rt.Options(4, 8, sample_begin=4 * p, accum_in=p > 0, accum_out=True, no_resolve=p < 2, resolve_spp=12,
                       sample_items=True)
        rt.render_device(h, o, W, H, out.data_ptr(), accum.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint8).reshape(H, W, 4), want)
    assert np.array_equal(accum.cpu().numpy(), want_acc)


def test_sample_items_are_chosen_for_heavy_scenes_only(gpu_rt, scenes):
    rt = gpu_rt
    _, st = _render(rt, rt.load_world(scenes.c5_world()), 64, 36, 4, 4)
    assert st.sample_items == 1                          # 10,000 primitives, few pixels per lane
    _, st = _render(rt, rt.load_world(scenes.default_world()), 64, 36, 4, 4)
    assert st.sample_items == 0                          # 8 spheres: the HBM round trip would not pay


def test_render_abi_call_is_16spp_depth8(gpu_rt, ob, scenes):
    """lib.rs:49-57: render(fb, handle) == Options::new(16, 8); the frame lands in the caller's
    buffer and the returned struct points at it."""
    rt = gpu_rt
    h = rt.load_world(scenes.example_world())
    fb = rt.Framebuffer(200, 200)                       # examples/c_raytracer.rs:50-51
    ret = rt.lib().render(fb._c(), h.ptr)
    assert rt.last_error() == ""
    assert (ret.width, ret.height, ret.pixels) == (200, 200, fb.pixels.ctypes.data)
    cam, world = ob.parse_input(scenes.example_world())
    want, _, _ = ob.ray_trace(world, cam, 200, 200, 16, 8)
    assert np.array_equal(fb.pixels, want)
    assert (fb.pixels[:, :, 3] == 255).all()            # alpha is always 255 (color.rs:21-23)


def test_plain_c_caller_of_the_reference_abi(gpu_rt, ob, scenes, tmp_path):
    """tests/c_caller/c_caller.c uses nothing but raytracer.h (load_world / move_camera_position /
    render) — the counterpart of examples/c_raytracer.rs: compiled with gcc against include/,
    linked with libraytracer.so, its frame equals the oracle's."""
    import subprocess
    lib = ROOT / "rust-swift-raytracer_b200" / "lib"
    exe = tmp_path / "c_caller"
    r = subprocess.run(["/usr/bin/gcc", "-std=c11", "-O1", "-Wall", "-Werror", f"-I{ROOT / 'include'}",
                        str(ROOT / "tests" / "c_caller" / "c_caller.c"), f"-L{lib}", "-lraytracer",
                        f"-Wl,-rpath,{lib}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    world_txt = tmp_path / "world.txt"
    world_txt.write_text(scenes.example_world())
    cam, world = ob.parse_input(scenes.example_world())
    for move in (None, (0.25, 0.125, -0.5)):
        out = tmp_path / "frame.rgba"
        args = [str(exe), str(world_txt), "200", "200", str(out)] + ([str(v) for v in move] if move else [])
        r = subprocess.run(args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.returncode, r.stderr)
        got = np.frombuffer(out.read_bytes(), dtype=np.uint8).reshape(200, 200, 4)
        c = ob.move_camera_position(cam, *move) if move else cam
        want, _, _ = ob.ray_trace(world, c, 200, 200, 16, 8)
        assert np.array_equal(got, want)


def test_camera_move_then_render(gpu_rt, ob, scenes):
    """GameView.swift:198-216: handle->camera = move_camera_position(handle->camera, ...)."""
    rt = gpu_rt
    h = rt.load_world(scenes.default_world())
    cam, world = ob.parse_input(scenes.default_world())
    for step in ((0.1, 0.0, 0.0), (0.0, 0.1, 0.0), (0.0, 0.0, -0.1)):
        rt.move_camera_position(h, *step)
        cam = ob.move_camera_position(cam, *step)
        got, _ = _render(rt, h, 96, 54, 2, 8)
        want, _, _ = ob.ray_trace(world, cam, 96, 54, 2, 8)
        assert np.array_equal(got, want)


def test_emission_and_all_materials_via_world_builder(gpu_rt, ob):
    """MaterialType::Emission is unreachable from the text parser (parser.rs:171-174)."""
    rt = gpu_rt
    h = rt.world_new((0, 0, 0), 1.5)
    w = ob.World()
    prims = [((0, 0, -1.2), 0.5, rt.EMISSION, (2.0, 1.5, 0.5), 0.0),
             ((1.1, 0, -1.2), 0.5, rt.METAL, (0.8, 0.8, 0.8), 0.4),
             ((-1.1, 0, -1.2), 0.5, rt.DIELECTRIC, (1, 1, 1), 1.5),
             ((0, -100.5, -1), 100.0, rt.DIFFUSE, (0.5, 0.6, 0.7), 0.0)]
    for c, r, m, col, p in prims:
        h.add_sphere(c, r, m, col, p)
        w.add_sphere(c, r, ob.material(m, col, p))
    tri = ((-2, -0.5, -3), (2, -0.5, -3), (0, 2.5, -3))
    h.add_triangle(*tri, rt.EMISSION, (0.2, 0.9, 0.2), 0.0)
    w.add_triangle(*tri, ob.material(ob.EMISSION, (0.2, 0.9, 0.2)))
    cam = ob.camera_new_at((0, 0, 0), 1.5)
    got, st = _render(rt, h, 120, 80, 8, 6)
    want, rays, _ = ob.ray_trace(w, cam, 120, 80, 8, 6)
    assert st.rays == rays and np.array_equal(got, want)
    # editing the world invalidates the cached device scene
    h.add_sphere((0, 0.9, -1.2), 0.3, rt.DIFFUSE, (0.9, 0.1, 0.1))
    w.add_sphere((0, 0.9, -1.2), 0.3, ob.material(ob.DIFFUSE, (0.9, 0.1, 0.1)))
    got, _ = _render(rt, h, 120, 80, 2, 6)
    want, _, _ = ob.ray_trace(w, cam, 120, 80, 2, 6)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("W,H,spp,depth", [(1, 1, 2, 4), (2, 2, 1, 8), (5, 3, 0, 8), (5, 3, 2, 0), (7, 1, 1, 3),
                                           (1, 9, 1, 3), (33, 5, 1, 1), (8, 4, 3, 2)])
def test_degenerate_frames(gpu_rt, ob, scenes, W, H, spp, depth):
    """W or H of 1 divides by zero (common.rs:335-336: NaN rays), spp 0 resolves 0/0, depth 0 is
    black: the kernel must agree with the oracle on every one."""
    rt = gpu_rt
    h = rt.load_world(scenes.example_world())
    cam, world = ob.parse_input(scenes.example_world())
    got, st = _render(rt, h, W, H, spp, depth)
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    assert st.rays == rays and np.array_equal(got, want)


def test_empty_world_is_sky(gpu_rt, ob):
    rt = gpu_rt
    src = "camera origin 0.0 0.0 0.0 aspect 1.5;"
    h = rt.load_world(src)
    cam, world = ob.parse_input(src)
    got, st = _render(rt, h, 64, 40, 3, 8)
    want, rays, _ = ob.ray_trace(world, cam, 64, 40, 3, 8)
    assert st.rays == rays == 64 * 40 * 3 and np.array_equal(got, want)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("RT_FUZZ_SEEDS", "24"))))
def test_random_worlds_fuzz(gpu_rt, ob, seed):
    """Awkward random geometry through every conservative filter (cases.random_world): bit-exact
    against the oracle's plain loops, pixel-item and sample-item scheduling alike."""
    rt = gpu_rt
    text = cases.random_world(seed)
    cam, world = ob.parse_input(text)
    W, H, spp, depth = 96, 64, 3, 8
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    h = rt.load_world(text)
    for items in (False, True):
        got, st = _render(rt, h, W, H, spp, depth, sample_items=items)
        assert st.filtered == 1 and st.rays == rays and np.array_equal(got, want), (seed, items)
    got, st = _render(rt, h, W, H, spp, depth, group_cull=True)             # opt-in acceleration mode
    assert st.culled == 1 and st.rays == rays and np.array_equal(got, want), seed


def test_seed_changes_the_realisation_and_matches_oracle(gpu_rt, ob, scenes):
    rt = gpu_rt
    h = rt.load_world(scenes.default_world())
    cam, world = ob.parse_input(scenes.default_world())
    a, _ = _render(rt, h, 80, 45, 2, 8, seed=12345)
    b, _ = _render(rt, h, 80, 45, 2, 8)
    want, _, _ = ob.ray_trace(world, cam, 80, 45, 2, 8, seed=12345)
    assert np.array_equal(a, want) and not np.array_equal(a, b)


def test_scene_too_large_for_shared_memory_uses_global_path(gpu_rt, ob, scenes):
    """> 227 KB of primitives: the kernel reads the list from global memory instead."""
    rt = gpu_rt
    text = scenes.synthetic_world(15000, 500, seed=77)
    h = rt.load_world(text)
    cam, world = ob.parse_input(text)
    got, st = _render(rt, h, 48, 27, 1, 6)
    want, rays, _ = ob.ray_trace(world, cam, 48, 27, 1, 6)
    assert st.resident == 0 and st.filtered == 1 and st.rays == rays and np.array_equal(got, want)


def test_shared_reciprocal_divide_equals_ieee_divide(gpu_rt):
    """rt_trace.cuh div3 / pixel_uv: 3 x 2^29 operand sets (moderate, extreme, zero, denormal,
    NaN/inf exponents), every quotient compared bit for bit with the compiler's `/`."""
    for seed in (1, 2, 3):
        assert gpu_rt.selftest_division(1 << 29, seed) == 0


def test_range_guarded_sqrt_equals_ieee_sqrt_for_every_float(gpu_rt):
    """rt_trace.cuh sqrt_ranged / sqrt_in_range: all 2^31 non-negative bit patterns (and the negatives' predicate)."""
    assert gpu_rt.selftest_sqrt() == 0


# ------------------------------------------------------------------ progressive, shards, device API

def test_progressive_passes_equal_one_pass(gpu_rt, ob, scenes):
    """16 = 4 x 4 spp through the float4 accumulator (device memory) == one 16-spp launch, and
    == the oracle's accumulator bit for bit."""
    import torch
    rt = gpu_rt
    W, H = 160, 90
    h = rt.load_world(scenes.default_world())
    cam, world = ob.parse_input(scenes.default_world())
    want, rays, want_acc = ob.ray_trace(world, cam, W, H, 16, 8, want_accum=True)
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    total = 0
    for p in range(4):
        st = rt.RenderStats()
        o = rt.Options(4, 8, sample_begin=4 * p, accum_in=p > 0, accum_out=True, no_resolve=p < 3, resolve_spp=16)
        rt.render_device(h, o, W, H, out.data_ptr(), accum.data_ptr(), 0, st)
        total += st.rays
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint8).reshape(H, W, 4)
    assert total == rays
    assert np.array_equal(got, want)
    assert np.array_equal(accum.cpu().numpy(), want_acc)


@pytest.mark.parametrize("count,tile_rows", [(2, 16), (3, 4), (8, 8)])
def test_row_tile_shards_reassemble_the_frame(gpu_rt, scenes, count, tile_rows):
    """Each shard renders only its tiles (host path scatters them into the caller's frame);
    the union over shards is the unsharded frame and the ray counts add up."""
    rt = gpu_rt
    W, H = 200, 117
    h = rt.load_world(scenes.example_world())
    full, st_full = _render(rt, h, W, H, 2, 8)
    fb = rt.Framebuffer(W, H)
    fb.pixels[...] = 0
    rays = 0
    for i in range(count):
        st = rt.RenderStats()
        rt.render_with_options(fb, h, rt.Options(2, 8, tile_rows=tile_rows, shard_index=i, shard_count=count), st)
        rays += st.rays
        rows = [r for r0, r1 in rt.shard_tiles(H, tile_rows, i, count) for r in range(r0, r1)]
        assert np.array_equal(fb.pixels[rows], full[rows])
    assert rays == st_full.rays and np.array_equal(fb.pixels, full)


def test_device_resident_render_on_a_torch_stream(gpu_rt, ob, scenes):
    import torch
    rt = gpu_rt
    W, H = 128, 72
    h = rt.load_world(scenes.default_world())
    cam, world = ob.parse_input(scenes.default_world())
    out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        rt.render_device(h, rt.Options(3, 8), W, H, out.data_ptr(), 0, s.cuda_stream)
    s.synchronize()
    want, _, _ = ob.ray_trace(world, cam, W, H, 3, 8)
    assert np.array_equal(out.cpu().numpy().view(np.uint8).reshape(H, W, 4), want)


def test_render_into_a_device_resident_framebuffer(gpu_rt, scenes):
    """render() with framebuffer.pixels pointing at device memory leaves the frame on the GPU."""
    import torch
    rt = gpu_rt
    W, H = 96, 54
    h = rt.load_world(scenes.default_world())
    want, _ = _render(rt, h, W, H, 16, 8)
    out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    rt.lib().render(rt._CFramebuffer(W, H, out.data_ptr()), h.ptr)
    assert rt.last_error() == ""
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint8).reshape(H, W, 4), want)


def test_pinned_and_pageable_destinations_agree(gpu_rt, scenes):
    rt = gpu_rt
    h = rt.load_world(scenes.default_world())
    a, _ = _render(rt, h, 320, 180, 2, 8, pinned=True)
    b, _ = _render(rt, h, 320, 180, 2, 8, pinned=False)
    assert np.array_equal(a, b)


def test_pageable_frames_never_mix_consecutive_frames(gpu_rt, scenes):
    """A pageable destination is filled tile by tile WHILE the kernel renders (per-tile completion flags, staging
    frame reused from call to call).  If a tile were handed over before all of its pixels had landed, the caller
    would see pixels of the PREVIOUS frame: alternate two very different views for 40 frames and compare every one
    with the frame a pinned destination receives (no staging, no flags)."""
    rt = gpu_rt
    W, H, spp, depth = 1920, 1080, 32, 8
    views = []
    for text, move in ((scenes.default_world(), (0.0, 0.0, 0.0)), (scenes.example_world(), (1.5, 0.75, 0.5))):
        h = rt.load_world(text)
        rt.move_camera_position(h, *move)
        fb = rt.Framebuffer(W, H, pinned=True)
        rt.render_with_options(fb, h, rt.Options(spp, depth))
        views.append((h, fb.pixels.copy()))
    assert not np.array_equal(views[0][1], views[1][1])
    fb = rt.Framebuffer(W, H, pinned=False)
    for i in range(40):
        h, want = views[i & 1]
        fb.pixels[...] = 0x5A
        rt.render_with_options(fb, h, rt.Options(spp, depth))
        assert np.array_equal(fb.pixels, want), f"frame {i}: {(fb.pixels != want).any(axis=2).sum()} pixels differ"


def test_interactive_progressive_frame(gpu_rt, ob, scenes):
    """rt_render_progressive: 5 calls of 3 spp == one 15-spp frame bit for bit; a
    camera move (GameView.swift:198-216) or a resize restarts the accumulation."""
    rt = gpu_rt
    W, H = 160, 90
    h = rt.load_world(scenes.default_world())
    cam, world = ob.parse_input(scenes.default_world())
    fb = rt.Framebuffer(W, H)
    for k in range(5):
        assert rt.render_progressive(fb, h, rt.Options(3, 8)) == 3 * (k + 1)
        want, _, _ = ob.ray_trace(world, cam, W, H, 3 * (k + 1), 8)
        assert np.array_equal(fb.pixels, want), k
    rt.move_camera_position(h, 0.25, 0.0, 0.0)
    cam = ob.move_camera_position(cam, 0.25, 0.0, 0.0)
    assert rt.render_progressive(fb, h, rt.Options(2, 8)) == 2
    want, _, _ = ob.ray_trace(world, cam, W, H, 2, 8)
    assert np.array_equal(fb.pixels, want)
    fb2 = rt.Framebuffer(W // 2, H // 2)
    assert rt.render_progressive(fb2, h, rt.Options(1, 8)) == 1
    assert rt.render_progressive(fb2, h, rt.Options(4, 8)) == 5
    want, _, _ = ob.ray_trace(world, cam, W // 2, H // 2, 5, 8)
    assert np.array_equal(fb2.pixels, want)


def test_concurrent_callers_are_serialised(gpu_rt, scenes):
    """The reference is single-threaded; the library guards its device state with one mutex, so
    threads rendering different worlds at the same time get the frames they would get alone."""
    import threading
    rt = gpu_rt
    jobs = [(scenes.default_world(), 160, 90, 4), (scenes.example_world(), 120, 120, 3), (scenes.c3_world(), 48, 27, 1)]
    alone = []
    for text, W, H, spp in jobs:
        got, _ = _render(rt, rt.load_world(text), W, H, spp, 8)
        alone.append(got)
    results = [None] * len(jobs)

    def work(i):
        text, W, H, spp = jobs[i]
        h = rt.load_world(text)
        for _ in range(5):
            fb = rt.Framebuffer(W, H)
            rt.lib().render_with_options(fb._c(), h.ptr, rt.Options(spp, 8)._c(None))
            results[i] = fb.pixels.copy()

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for got, want in zip(results, alone):
        assert np.array_equal(got, want)


def test_errors_do_not_cross_the_abi(gpu_rt, scenes):
    rt = gpu_rt
    h = rt.load_world(scenes.default_world())
    fb = rt.Framebuffer(16, 8)
    with pytest.raises(rt.RenderError, match="tile_rows"):
        rt.render_with_options(fb, h, rt.Options(1, 1, tile_rows=6))
    with pytest.raises(rt.RenderError, match="shard"):
        rt.render_with_options(fb, h, rt.Options(1, 1, shard_index=2, shard_count=2))
    with pytest.raises(rt.ParseError):
        rt.load_world("sphere center 0 0 0 radius 1 material X;")


# ------------------------------------------------------------------ statistical agreement

def _noise_floor(ob, world, cam, W, H, spp, depth):
    a, _, _ = ob.ray_trace(world, cam, W, H, spp, depth, seed=1)
    b, _, _ = ob.ray_trace(world, cam, W, H, spp, depth, seed=2)
    return _rmse(a, b)


def test_stochastic_render_agrees_with_the_reference_serial_stream(gpu_rt, ob, scenes):
    """The reference draws ONE serial xorshift stream per frame (common.rs:321), so its image is
    a different realisation of the same estimator.  Stated tolerance: RMSE(GPU, serial oracle)
    <= 1.25 x RMSE of two independent oracle renders at the same spp (both are differences of
    two independent N-spp estimates), and mean colour within 0.5/255."""
    rt = gpu_rt
    W, H, spp = 400, 224, 50          # BASELINE config 1 as src/main.rs renders it (main.rs:86-99)
    h = cases.product_scene(rt, scenes, "default", cases.LOOK_AT_CLI)
    cam, world = cases.oracle_scene(ob, scenes, "default", cases.LOOK_AT_CLI)
    got, _ = _render(rt, h, W, H, spp, 8)
    serial, _, _ = ob.ray_trace(world, cam, W, H, spp, 8, rng_mode=ob.RNG_SERIAL)
    floor = _noise_floor(ob, world, cam, W, H, spp, 8)
    r = _rmse(got, serial)
    assert r <= 1.25 * floor, (r, floor)
    assert abs(got[:, :, :3].mean() - serial[:, :, :3].mean()) < 0.5


@pytest.mark.parametrize("key,W,H,spp,depth", [("default", 400, 224, 50, 8), ("example", 200, 200, 16, 8),
                                               ("c3", 192, 108, 8, 8)])
def test_fast_math_kernel_is_statistically_equivalent(gpu_rt, ob, scenes, key, W, H, spp, depth):
    """RT_OPT_FAST_MATH (FMA, rsqrt/rcp approximations): same seeds, so almost every path is
    identical and the rest flip at silhouettes.  Stated tolerance: RMSE <= 0.35 x the noise
    floor of two independent renders, <= 1 % of pixels differ by more than 1/255, ray count
    within 0.1 %."""
    rt = gpu_rt
    h = rt.load_world(cases.scene_text(scenes, key))
    cam, world = ob.parse_input(cases.scene_text(scenes, key))
    exact, st_e = _render(rt, h, W, H, spp, depth)
    fast, st_f = _render(rt, h, W, H, spp, depth, fast_math=True)
    floor = _noise_floor(ob, world, cam, W, H, spp, depth)
    d = np.abs(exact.astype(int) - fast.astype(int)).max(axis=2)
    assert _rmse(exact, fast) <= 0.35 * floor, (_rmse(exact, fast), floor)
    assert (d > 1).mean() <= 0.01
    assert abs(int(st_e.rays) - int(st_f.rays)) <= 1e-3 * st_e.rays


# ------------------------------------------------------------------ full-size properties

def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_c2_full_size_properties(gpu_rt, ob, scenes):
    """BASELINE config 2 (default scene, 1920x1080, 64 spp, depth 8) is far beyond the oracle's
    reach (~400 M ray segments), so it is checked through properties:
      idempotence (same seed -> same bytes), shard union == whole frame, 4 x 16 spp progressive
      == one 64-spp launch, sample-count bookkeeping, alpha == 255, and the oracle at 1 spp."""
    import torch
    rt = gpu_rt
    W, H, spp, depth = 1920, 1080, 64, 8
    h = rt.load_world(scenes.default_world())
    a, st = _render(rt, h, W, H, spp, depth, pinned=True)
    b, st2 = _render(rt, h, W, H, spp, depth, pinned=True)
    assert _sha(a) == _sha(b) and st.rays == st2.rays
    assert st.samples == W * H * spp and W * H * spp <= st.rays <= W * H * spp * depth
    assert (a[:, :, 3] == 255).all()
    # shards
    fb = rt.Framebuffer(W, H, pinned=True)
    rays = 0
    for i in range(4):
        s = rt.RenderStats()
        rt.render_with_options(fb, h, rt.Options(spp, depth, shard_index=i, shard_count=4), s)
        rays += s.rays
    assert rays == st.rays and _sha(fb.pixels) == _sha(a)
    # progressive
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    for p in range(4):
        o = rt.Options(16, depth, sample_begin=16 * p, accum_in=p > 0, accum_out=True, no_resolve=p < 3, resolve_spp=spp)
        rt.render_device(h, o, W, H, out.data_ptr(), accum.data_ptr(), 0)
    torch.cuda.synchronize()
    assert _sha(out.cpu().numpy()) == _sha(a)
    # sample-item scheduling at a size where the 1 GiB sample buffer forces two chunks (2 x 32 spp)
    c, st3 = _render(rt, h, W, H, spp, depth, pinned=True, sample_items=True)
    assert st3.sample_items == 1 and st3.launches == 4 and st3.rays == st.rays and _sha(c) == _sha(a)
    # oracle at full resolution, 1 spp (the per-sample RNG makes every sample independent)
    cam, world = ob.parse_input(scenes.default_world())
    small, _, _ = ob.ray_trace(world, cam, W, H, 1, depth)
    one, _ = _render(rt, h, W, H, 1, depth)
    assert np.array_equal(one, small)


def test_c2_full_size_equals_oracle(gpu_rt, ob, scenes):
    """The whole BASELINE config-2 frame (1920x1080, 64 spp, depth 8; ~381 M ray segments)
    against the OpenMP oracle: bit-exact pixels and identical ray count.  ~10 s on 8 cores."""
    rt = gpu_rt
    W, H, spp, depth = 1920, 1080, 64, 8
    h = rt.load_world(scenes.default_world())
    got, st = _render(rt, h, W, H, spp, depth, pinned=True)
    cam, world = ob.parse_input(scenes.default_world())
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    assert st.rays == rays
    assert np.array_equal(got, want)


def test_c3_c5_reduced_spp_full_resolution_bands(gpu_rt, ob, scenes):
    """Configs 3 and 5 at reduced size still run the 1,000- / 10,000-primitive scenes
    bit-exactly (shared-memory staged list; triangles through the Mesh path)."""
    rt = gpu_rt
    for key, W, H, spp, depth in (("c3", 480, 270, 2, 8), ("c5", 96, 54, 2, 16)):
        text = cases.scene_text(scenes, key)
        h = rt.load_world(text)
        cam, world = ob.parse_input(text)
        got, st = _render(rt, h, W, H, spp, depth)
        want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
        assert st.resident == 1 and st.filtered == 1 and st.rays == rays and np.array_equal(got, want), key


def test_c1_full_config_equals_oracle(gpu_rt, ob, scenes):
    """BASELINE config 1 exactly as src/main.rs renders it (main.rs:86-99): world.txt, new_look_at camera,
    400x224, 50 spp, depth 8 — bit-exact pixels and ray count against the oracle."""
    rt = gpu_rt
    W, H, spp, depth = 400, 224, 50, 8
    h = cases.product_scene(rt, scenes, "default", cases.LOOK_AT_CLI)
    cam, world = cases.oracle_scene(ob, scenes, "default", cases.LOOK_AT_CLI)
    got, st = _render(rt, h, W, H, spp, depth)
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    assert st.rays == rays and np.array_equal(got, want)


def test_c3_full_config_band_equals_oracle(gpu_rt, ob, scenes):
    """BASELINE config 3 at its REAL size — 1,000 spheres, 1920x1080, 256 spp, depth 8: the whole frame on the
    GPU (sample items, 8 chunks of 32 spp under the 1 GiB sample-buffer cap), and one full-width 16-row tile
    of it against the oracle (per-sample RNG: a band of the frame is exactly that band).  Bit-exact."""
    rt = gpu_rt
    W, H, spp, depth = 1920, 1080, 256, 8
    text = scenes.c3_world()
    h = rt.load_world(text)
    got, st = _render(rt, h, W, H, spp, depth, pinned=True)
    assert st.sample_items == 1 and st.launches == 16 and st.filtered == 1 and st.samples == W * H * spp
    cam, world = ob.parse_input(text)
    r0, r1 = 640, 656                                     # tile 40: small spheres, ground and sky reflections
    want, rays_band, _ = ob.ray_trace(world, cam, W, H, spp, depth, rows=(r0, r1))
    assert np.array_equal(got[r0:r1], want[r0:r1])
    # the same tile alone, through the shard interface: identical pixels, and its ray count is the oracle's
    fb = rt.Framebuffer(W, H)
    s1 = rt.RenderStats()
    rt.render_with_options(fb, h, rt.Options(spp, depth, tile_rows=16, shard_index=40, shard_count=68), s1)
    assert s1.rays == rays_band and np.array_equal(fb.pixels[r0:r1], want[r0:r1])


def test_c5_full_config_band_equals_oracle(gpu_rt, ob, scenes):
    """BASELINE config 5 at its REAL size — 8,000 spheres + 2,000 triangles, 1280x720, 16 spp, depth 16:
    whole frame on the GPU, one full-width 16-row tile against the oracle.  Bit-exact."""
    rt = gpu_rt
    W, H, spp, depth = 1280, 720, 16, 16
    text = scenes.c5_world()
    h = rt.load_world(text)
    got, st = _render(rt, h, W, H, spp, depth, pinned=True)
    assert st.resident == 1 and st.filtered == 1 and st.samples == W * H * spp
    cam, world = ob.parse_input(text)
    r0, r1 = 400, 416
    want, rays_band, _ = ob.ray_trace(world, cam, W, H, spp, depth, rows=(r0, r1))
    assert np.array_equal(got[r0:r1], want[r0:r1])
    fb = rt.Framebuffer(W, H)
    s1 = rt.RenderStats()
    rt.render_with_options(fb, h, rt.Options(spp, depth, tile_rows=16, shard_index=25, shard_count=45), s1)
    assert s1.rays == rays_band and np.array_equal(fb.pixels[r0:r1], want[r0:r1])


def test_c4_geometry_16_fused_passes_equal_oracle(gpu_rt, ob, scenes):
    """BASELINE config 4's shape — 3840x2160, 16 progressive passes — at 1 spp per pass: ONE persistent launch
    traces all 16 passes (pass-major queue, sums through HBM), and the frame is the oracle's 16-spp frame."""
    rt = gpu_rt
    W, H, depth = 3840, 2160, 8
    h = rt.load_world(scenes.default_world())
    got, st = _render(rt, h, W, H, 16, depth, pinned=True, passes=16)
    assert st.launches == 1 and st.passes_fused == 16
    cam, world = ob.parse_input(scenes.default_world())
    want, rays, _ = ob.ray_trace(world, cam, W, H, 16, depth)
    assert st.rays == rays and np.array_equal(got, want)


@pytest.mark.parametrize("key,W,H,spp,passes,depth", [("default", 333, 170, 12, 4, 8), ("example", 200, 120, 8, 8, 8),
                                                      ("default", 64, 36, 6, 2, 1), ("c3", 96, 54, 4, 2, 8)])
def test_fused_passes_equal_single_pass_and_oracle_sums(gpu_rt, ob, scenes, key, W, H, spp, passes, depth):
    """RtRenderOptions.passes: k passes in one launch == one pass of the total, bit for bit — frame, ray count
    and (accum_out) the float4 sums; RT_OPT_RESOLVE_EACH_PASS changes nothing in the final frame; shards too."""
    import torch
    rt = gpu_rt
    text = cases.scene_text(scenes, key)
    h = rt.load_world(text)
    cam, world = ob.parse_input(text)
    want, rays, want_acc = ob.ray_trace(world, cam, W, H, spp, depth, want_accum=True)
    got, st = _render(rt, h, W, H, spp, depth, passes=passes, sample_items=False)
    assert st.passes_fused == passes and st.launches == 1 and st.rays == rays and np.array_equal(got, want)
    got, st = _render(rt, h, W, H, spp, depth, passes=passes, resolve_each_pass=True, sample_items=False)
    assert st.passes_fused == passes and np.array_equal(got, want)
    accum = torch.full((H, W, 4), 7.0, dtype=torch.float32, device="cuda")       # stale contents must not matter
    out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    s2 = rt.RenderStats()
    rt.render_device(h, rt.Options(spp, depth, passes=passes, accum_out=True, sample_items=False), W, H, out.data_ptr(),
                     accum.data_ptr(), 0, s2)
    assert s2.passes_fused == passes and s2.rays == rays
    assert np.array_equal(out.cpu().numpy().view(np.uint8).reshape(H, W, 4), want)
    assert np.array_equal(accum.cpu().numpy(), want_acc)
    fb = rt.Framebuffer(W, H)
    for i in range(3):
        rt.render_with_options(fb, h, rt.Options(spp, depth, passes=passes, shard_index=i, shard_count=3, tile_rows=8,
                                                 sample_items=False))
    assert np.array_equal(fb.pixels, want)
    # sample items / continuing an accumulator cannot be fused: the total is traced as one pass, same bits
    got, st = _render(rt, h, W, H, spp, depth, passes=passes, sample_items=True)
    assert st.passes_fused == 0 and st.sample_items == 1 and np.array_equal(got, want)


def _blocks(rt, W, H, n):
    import torch
    b = [torch.zeros(rt.shard_block_bytes(W, H), dtype=torch.uint8, device="cuda") for _ in range(n)]
    torch.cuda.synchronize()           # the library renders on its own stream
    return b


@pytest.mark.parametrize("passes", [1, 4])
def test_work_stealing_queues_on_one_gpu(gpu_rt, ob, scenes, passes):
    """Cross-GPU work stealing, exercised deterministically on ONE device: three shard blocks whose owners
    never start (their counters stay 0 = 'everything unassigned'); the launch of shard 0 renders its own
    tiles, then raids queues 2 and 1 until the whole frame is done — every stolen pixel-pass reads and
    writes the sums in its victim's block.  Frame and ray count equal the oracle's."""
    import torch
    rt = gpu_rt
    W, H, spp, depth = 200, 117, 8, 8
    text = scenes.example_world()
    h = rt.load_world(text)
    cam, world = ob.parse_input(text)
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    blocks = _blocks(rt, W, H, 3)
    out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    st = rt.RenderStats()
    o = rt.Options(spp, depth, passes=passes, tile_rows=8, shard_index=0, shard_count=3, full_frame_out=True,
                   peer_queues=[(blocks[0].data_ptr(), 0), (blocks[2].data_ptr(), 2), (blocks[1].data_ptr(), 1)])
    rt.render_device(h, o, W, H, out.data_ptr(), 0, 0, st)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint8).reshape(H, W, 4), want)
    assert st.rays == rays and st.passes_fused == (passes if passes > 1 else 0)
    per_tile = ((W + 7) // 8) * 2 * 32                       # slots of one 8-row tile
    tiles = [len(rt.shard_tiles(H, 8, i, 3)) for i in range(3)]
    assert st.stolen_slots == (tiles[1] + tiles[2]) * per_tile * passes
    # afterwards every queue reads 'empty': a second launch of another shard finds nothing to steal
    out2 = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    o2 = rt.Options(spp, depth, passes=passes, tile_rows=8, shard_index=1, shard_count=3, full_frame_out=True,
                     peer_queues=[(blocks[1].data_ptr(), 1), (blocks[2].data_ptr(), 2), (blocks[0].data_ptr(), 0)])
    s2 = rt.RenderStats()
    rt.render_device(h, o2, W, H, out2.data_ptr(), 0, 0, s2)
    torch.cuda.synchronize()
    rows = [r for r0, r1 in rt.shard_tiles(H, 8, 1, 3) for r in range(r0, r1)]
    got2 = out2.cpu().numpy().view(np.uint8).reshape(H, W, 4)
    assert s2.stolen_slots == 0 and np.array_equal(got2[rows], want[rows])
    other = [r for r in range(H) if r not in rows]
    assert not got2[other].any()


@pytest.mark.parametrize("passes", [1, 3])
def test_row_gather_stages_locally_and_copies_vectorised(gpu_rt, ob, scenes, passes):
    """RT_OPT_ROW_GATHER (what ranks != 0 use for a frame in rank 0's memory), on one device: the launch renders into a
    local, zeroed frame and a second kernel moves every pixel found there into the destination as 16-byte vectors —
    its own tiles and whatever it stole — skipping zero words (pixels somebody else traced), so nothing already in the
    destination is wiped.  Ragged frame: 201 x 117, 8-row tiles (last tile 5 rows, rows of 804 bytes)."""
    import torch
    rt = gpu_rt
    W, H, spp, depth = 201, 117, 6, 8
    text = scenes.example_world()
    h = rt.load_world(text)
    cam, world = ob.parse_input(text)
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    # (a) three shards, each launched alone with row gather: together they fill the frame
    out = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    total = 0
    for i in range(3):
        st = rt.RenderStats()
        rt.render_device(h, rt.Options(spp, depth, passes=passes, tile_rows=8, shard_index=i, shard_count=3,
                                       full_frame_out=True, row_gather=True), W, H, out.data_ptr(), 0, 0, st)
        total += st.rays
    torch.cuda.synchronize()
    assert total == rays and np.array_equal(out.cpu().numpy().view(np.uint8).reshape(H, W, 4), want)
    # (b) shard 1 with row gather raids the queues of shards 2 and 0, whose owners never start: own and stolen pixels
    #     all arrive by the vector copy
    blocks = _blocks(rt, W, H, 3)
    out = torch.full((H, W), 0x01020304, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    st = rt.RenderStats()
    rt.render_device(h, rt.Options(spp, depth, passes=passes, tile_rows=8, shard_index=1, shard_count=3, full_frame_out=True,
                                   row_gather=True, peer_queues=[(blocks[1].data_ptr(), 1), (blocks[2].data_ptr(), 2),
                                                                 (blocks[0].data_ptr(), 0)]), W, H, out.data_ptr(), 0, 0, st)
    torch.cuda.synchronize()
    assert st.rays == rays and st.stolen_slots > 0
    assert np.array_equal(out.cpu().numpy().view(np.uint8).reshape(H, W, 4), want)


def test_work_stealing_rejects_bad_queue_tables(gpu_rt, scenes):
    import torch
    rt = gpu_rt
    h = rt.load_world(scenes.default_world())
    blocks = _blocks(rt, 64, 32, 2)
    out = torch.zeros((32, 64), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    q = [(blocks[0].data_ptr(), 0), (blocks[1].data_ptr(), 1)]
    with pytest.raises(rt.RenderError, match="full-frame"):
        rt.render_device(h, rt.Options(1, 1, shard_index=0, shard_count=2, peer_queues=q), 64, 32, out.data_ptr(), 0, 0)
    with pytest.raises(rt.RenderError, match="every shard"):
        rt.render_device(h, rt.Options(1, 1, shard_index=0, shard_count=3, full_frame_out=True, peer_queues=q), 64, 32,
                         out.data_ptr(), 0, 0)
    with pytest.raises(rt.RenderError, match="own block"):
        rt.render_device(h, rt.Options(1, 1, shard_index=0, shard_count=2, full_frame_out=True,
                                       peer_queues=[(blocks[1].data_ptr(), 1), (blocks[1].data_ptr(), 1)]), 64, 32,
                         out.data_ptr(), 0, 0)


def test_progressive_frame_on_another_device(gpu_rt, ob, scenes):
    """rt_render_progressive with options->device != the caller's current device: the sums live on the device
    that renders, and the caller's current device is left alone."""
    import torch
    rt = gpu_rt
    if rt.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 96, 54
    h = rt.load_world(scenes.default_world())
    cam, world = ob.parse_input(scenes.default_world())
    torch.cuda.set_device(0)
    fb = rt.Framebuffer(W, H)
    assert rt.render_progressive(fb, h, rt.Options(2, 8, device=1)) == 2
    assert rt.render_progressive(fb, h, rt.Options(3, 8, device=1)) == 5
    want, _, _ = ob.ray_trace(world, cam, W, H, 5, 8)
    assert np.array_equal(fb.pixels, want) and torch.cuda.current_device() == 0


def test_unbalanced_shards_are_rebalanced_by_stealing(gpu_rt, scenes):
    """Two devices, two tiles: the upper half of the frame (sky, one segment per sample) and the lower half (the
    ground, long paths).  The device that owns the cheap tile finishes early and raids the other one's queue."""
    rt = gpu_rt
    if rt.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 1024, 512
    h = rt.load_world(scenes.default_world())
    full, st1 = _render(rt, h, W, H, 32, 8)
    for passes in (1, 4):
        got, st = _render(rt, h, W, H, 32, 8, n_devices=2, tile_rows=256, passes=passes)
        assert st.devices == 2 and st.peer_gather == 1 and st.rays == st1.rays and np.array_equal(got, full)
        assert st.stolen_slots > 0, "the device with the sky tile should have raided the ground tile's queue"


def test_one_process_many_gpus_peer_store_gather(gpu_rt, ob, scenes):
    """render_with_options(n_devices=N): device d renders tiles d, d+N, ... and stores them
    straight into device 0's frame (peer mapping); the frame equals the single-GPU frame."""
    rt = gpu_rt
    n = rt.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 640, 360
    h = rt.load_world(scenes.example_world())
    full, st1 = _render(rt, h, W, H, 4, 8)
    for nd in sorted({2, n}):
        for pinned in (False, True):
            got, st = _render(rt, h, W, H, 4, 8, pinned=pinned, n_devices=nd)
            assert np.array_equal(got, full), (nd, pinned)
            # one render kernel per device + one row-gather copy kernel on every device but the one that owns the frame
            assert st.devices == nd and st.rays == st1.rays and st.launches == 2 * nd - 1 and st.peer_gather == 1


def _dist_worker(rank, world, port, gather, q):
    import importlib
    import os
    import sys
    expect = gather
    if gather == "peer-unavailable":            # CUDA IPC fails on the non-zero ranks: everybody falls back
        os.environ["RT_DISABLE_IPC"] = "1"
        gather, expect = "peer", "nccl"
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=90))     # a dead peer must not hang the suite
    rt = importlib.import_module("rust-swift-raytracer_b200")
    multi = importlib.import_module("rust-swift-raytracer_b200.multi")
    scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")
    h = rt.load_world(scenes.example_world())
    W, H = 333, 170                                  # ragged: last tile is partial, tiles % world != 0
    r = multi.ShardedRenderer(rt, h, W, H, rank, world, tile_rows=16, gather=gather)
    assert r.gather == expect, (r.gather, expect, r.peer_error)
    assert r.steal == (expect == "peer")                         # shard blocks exchanged over CUDA IPC
    frame, rays = r.render(8, 8, passes=2, to_host=True, count_rays=True)
    if frame is not None:
        frame = frame.clone()                                    # rank 0: the host frame is reused by the next render()
    frame2, _ = r.render(8, 8, passes=1, to_host=True)          # single pass == two progressive passes
    t = torch.tensor([rays], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        q.put((frame.numpy().copy(), frame2.numpy().copy(), int(t.item())))
    r.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gather", ["peer", "nccl", "peer-unavailable"])
def test_one_process_per_gpu_sharded_frame_equals_single_gpu(gpu_rt, scenes, gather):
    """The bench's N > 1 path (one process per GPU, torch.distributed/NCCL): tile shards +
    gather (peer stores through CUDA IPC, or dist.gather) == the single-GPU frame, bit for bit."""
    import socket
    import torch.multiprocessing as mp
    rt = gpu_rt
    if rt.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, world, port, gather, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        frame, frame2, rays = q.get(timeout=180)
    finally:
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0
    h = rt.load_world(scenes.example_world())
    want, st = _render(rt, h, 333, 170, 8, 8)
    assert rays == st.rays
    assert np.array_equal(frame.view(np.uint8).reshape(170, 333, 4), want)
    assert np.array_equal(frame2, frame)


@pytest.mark.parametrize("key,W,H,spp,depth", [("c3", 480, 270, 2, 8), ("c5", 96, 54, 2, 16), ("c5mini", 160, 90, 2, 16)])
def test_group_cull_mode_is_bit_identical(gpu_rt, ob, scenes, key, W, H, spp, depth):
    """RT_OPT_GROUP_CULL (opt-in): bounding spheres over spatially ordered groups of
    8 spheres in front of the filter.  Same pixels, same ray counts, both kernels' sample scheduling;
    worlds below 64 spheres ignore the flag."""
    rt = gpu_rt
    text = cases.scene_text(scenes, key)
    h = rt.load_world(text)
    cam, world = ob.parse_input(text)
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    for items in (False, True):
        got, st = _render(rt, h, W, H, spp, depth, group_cull=True, sample_items=items)
        assert st.culled == 1 and st.rays == rays and np.array_equal(got, want), (key, items)
    fast_plain, sp = _render(rt, h, W, H, spp, depth, fast_math=True)
    fast_cull, sc = _render(rt, h, W, H, spp, depth, fast_math=True, group_cull=True)
    assert sc.culled == 1 and sc.rays == sp.rays and np.array_equal(fast_cull, fast_plain)   # same test decides
    small = rt.load_world(scenes.default_world())
    _, st = _render(rt, small, 64, 36, 2, 4, group_cull=True)
    assert st.culled == 0


def test_multi_gpu_tile_gather_when_two_devices(gpu_rt, scenes):
    """Two devices in one process: device 1 renders its shard, the tiles land in the frame."""
    rt = gpu_rt
    if rt.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 256, 144
    h = rt.load_world(scenes.default_world())
    full, _ = _render(rt, h, W, H, 2, 8)
    fb = rt.Framebuffer(W, H)
    for i in range(2):
        rt.render_with_options(fb, h, rt.Options(2, 8, shard_index=i, shard_count=2, device=i))
    assert np.array_equal(fb.pixels, full)


def test_two_paths_per_lane_kernels_are_bit_identical(scenes, tmp_path):
    """RT_PATHS_PER_LANE=2 (read once per process) selects the FILTER kernels that carry two paths per lane and test
    both rays against every primitive pair they load: same frames, same ray counts, pixel and sample items alike."""
    import subprocess
    import sys
    code = r'''
import importlib, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle_binding as ob
rt = importlib.import_module("rust-swift-raytracer_b200"); scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")
for text, W, H, spp, depth in ((scenes.c3_world(), 160, 90, 3, 8), (scenes.synthetic_world(800, 200, seed=10000), 96, 54, 2, 16)):
    cam, world = ob.parse_input(text)
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    h = rt.load_world(text)
    for items in (False, True):
        fb = rt.Framebuffer(W, H); st = rt.RenderStats()
        rt.render_with_options(fb, h, rt.Options(spp, depth, sample_items=items), st)
        assert st.paths_per_lane == 2 and st.filtered == 1, (st.paths_per_lane, st.filtered)
        assert st.rays == rays and np.array_equal(fb.pixels, want), items
print("ok")
''' % (str(ROOT), str(ROOT / "oracle"))
    import os
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, RT_PATHS_PER_LANE="2"))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
