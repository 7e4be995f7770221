"""Host-logic check of the DEVICE code: csrc/rt_trace.cuh (exact policy) compiled with g++ and
driven by a plain per-pixel loop (tests/hostsim) must equal the oracle bit for bit.  This is a
test-only build — the shipped library contains no CPU render path."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases

HERE = Path(__file__).resolve().parent


def _load(name):
    subprocess.run(["make", "-C", str(HERE / "hostsim")], check=True, capture_output=True)
    L = C.CDLL(str(HERE / "hostsim" / "build" / name))
    L.hostsim_render.restype = C.c_int
    L.hostsim_render.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.c_int32, C.c_int32,
                                 C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_uint64)]
    L.hostsim_divisor_mismatches.restype = C.c_uint64
    L.hostsim_divisor_mismatches.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_uint32]
    L.hostsim_decode_violations.restype = C.c_uint64
    L.hostsim_decode_violations.argtypes = [C.c_uint32] * 5 + [C.c_int32, C.c_int, C.c_int]
    L.hostsim_check_cull.restype = C.c_int
    L.hostsim_check_cull.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    return L


@pytest.fixture(scope="module")
def hostsim():
    return _load("libhostsim.so")


@pytest.fixture(scope="module")
def hostsim_perturbed():
    """The same build with rcp_approx (MUFU.RCP on the device) moved one ulp up or down pseudo-randomly."""
    return _load("libhostsim_perturb.so")


def test_cull_block_invariants(hostsim, scenes):
    """Block C of the scene blob (group-cull mode): every sphere exactly once, member records are the
    list's own, and every group's bound constants dominate what the derivation in rt_trace.cuh needs."""
    for text, want_groups in ((scenes.c3_world(), True), (scenes.c5_world(), True), (scenes.default_world(), False),
                              (cases.random_world(3, n_spheres=300), True), (cases.random_world(4, n_spheres=64), True)):
        ng, na = C.c_uint32(), C.c_uint32()
        assert hostsim.hostsim_check_cull(text.encode(), C.byref(ng), C.byref(na)) == 0
        assert (ng.value > 0) == want_groups and ng.value % 32 == 0
    # the C3 scene's ground sphere (r = 1000) sits in a group of its own that always passes
    ng, na = C.c_uint32(), C.c_uint32()
    assert hostsim.hostsim_check_cull(scenes.c3_world().encode(), C.byref(ng), C.byref(na)) == 0 and na.value == 1


@pytest.mark.parametrize("case", cases.SMALL_CASES, ids=[c[0] for c in cases.SMALL_CASES])
def test_device_header_on_host_equals_oracle(hostsim, ob, scenes, case):
    name, key, camera, W, H, spp, depth, fixed = case
    if W * H * spp > 200_000:
        W, H = W // 2, H // 2
    cam, world = cases.oracle_scene(ob, scenes, key, camera)
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth, fixed_jitter=fixed)
    out = np.zeros((H, W, 4), np.uint8)
    n = C.c_uint64()
    cf = cam.floats()
    rc = hostsim.hostsim_render(cases.scene_text(scenes, key).encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), W, H,
                                spp, depth, ob.SEED_DEFAULT, 1 if fixed else 0, 0, 0, out.ctypes.data, C.byref(n))
    assert rc == 0
    assert n.value == rays
    assert np.array_equal(out, want)


@pytest.mark.parametrize("key,W,H,spp,depth", [("c3", 64, 36, 2, 8), ("c5mini", 48, 27, 2, 16), ("c5", 24, 14, 1, 16)])
def test_cull_variant_equals_oracle(hostsim, ob, scenes, key, W, H, spp, depth):
    """RT_SPH_CULL: bounding spheres over spatially ordered groups of 8 in front of the filter —
    same pixels, same ray counts (the list-order tie-break is carried by the stored list index)."""
    cam, world = cases.oracle_scene(ob, scenes, key, "file")
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    out = np.zeros((H, W, 4), np.uint8)
    n = C.c_uint64()
    cf = cam.floats()
    rc = hostsim.hostsim_render(cases.scene_text(scenes, key).encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), W, H,
                                spp, depth, ob.SEED_DEFAULT, 0x40000000, 0, 0, out.ctypes.data, C.byref(n))
    assert rc == 0 and n.value == rays
    assert np.array_equal(out, want)


@pytest.mark.parametrize("key,W,H,spp,depth", [("c3", 64, 36, 2, 8), ("c5mini", 48, 27, 2, 16), ("example", 80, 80, 3, 8)])
def test_filter_variant_of_the_exact_policy_equals_oracle(hostsim, ob, scenes, key, W, H, spp, depth):
    """The conservative FMA sphere filter (block B, sphere_filter_group) in front of the exact
    test must not change a single bit or ray count."""
    cam, world = cases.oracle_scene(ob, scenes, key, "file")
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    out = np.zeros((H, W, 4), np.uint8)
    n = C.c_uint64()
    cf = cam.floats()
    rc = hostsim.hostsim_render(cases.scene_text(scenes, key).encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), W, H,
                                spp, depth, ob.SEED_DEFAULT, 0x80000000, 0, 0, out.ctypes.data, C.byref(n))
    assert rc == 0 and n.value == rays
    assert np.array_equal(out, want)
    # two paths per lane (closest_hit_n: both rays against every sphere / plane the lane loads), odd pixel count too
    for w2, h2 in ((W, H), (W - 1, H - 1)):
        want2, rays2, _ = ob.ray_trace(world, cam, w2, h2, spp, depth)
        out = np.zeros((h2, w2, 4), np.uint8)
        rc = hostsim.hostsim_render(cases.scene_text(scenes, key).encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), w2, h2,
                                    spp, depth, ob.SEED_DEFAULT, 0x20000000, 0, 0, out.ctypes.data, C.byref(n))
        assert rc == 0 and n.value == rays2
        assert np.array_equal(out, want2)


@pytest.mark.parametrize("seed", range(12))
def test_random_worlds_through_the_filters(hostsim, ob, seed):
    """Fuzz: awkward geometry (six decades of radii, far-away / huge / camera-containing spheres,
    needle / degenerate / distant triangles) through the conservative sphere filter, the triangle
    plane prefilter and the barycentric edge reject — every pixel and the ray count must equal
    the oracle's plain loops."""
    text = cases.random_world(seed)
    cam, world = ob.parse_input(text)
    W, H, spp, depth = 40, 28, 2, 6
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    cf = cam.floats()
    for flags in (0, 0x80000000, 0x40000000, 0x20000000):   # direct loops, FILTER, CULL, FILTER with two paths per lane
        out = np.zeros((H, W, 4), np.uint8)
        n = C.c_uint64()
        rc = hostsim.hostsim_render(text.encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), W, H, spp, depth,
                                    ob.SEED_DEFAULT, flags, 0, 0, out.ctypes.data, C.byref(n))
        assert rc == 0 and n.value == rays, (seed, flags)
        assert np.array_equal(out, want), (seed, flags)


@pytest.mark.parametrize("W,H,spp,depth", [(1, 1, 2, 4), (2, 2, 1, 8), (5, 3, 0, 8), (5, 3, 2, 0), (7, 1, 1, 3), (1, 9, 1, 3)])
def test_degenerate_frames(hostsim, ob, scenes, W, H, spp, depth):
    """W or H of 1 divides by zero in common.rs:335-336 (NaN rays -> black), spp 0 resolves 0/0,
    depth 0 returns black: all must agree with the oracle."""
    cam, world = ob.parse_input(scenes.example_world())
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    out = np.zeros((H, W, 4), np.uint8)
    n = C.c_uint64()
    cf = cam.floats()
    rc = hostsim.hostsim_render(scenes.example_world().encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), W, H, spp,
                                depth, ob.SEED_DEFAULT, 0, 0, spp, out.ctypes.data, C.byref(n))
    assert rc == 0 and n.value == rays
    assert np.array_equal(out, want)


@pytest.mark.parametrize("seed", range(40, 70))
def test_random_worlds_with_perturbed_reciprocals(hostsim_perturbed, ob, seed):
    """The conservative triangle filters (approximate quotient ta = num * rcp.approx(den), window margin 2^-19, edge
    reject on p(ta)) must not depend on the last bits of the approximation: with every rcp_approx result moved one
    ulp up or down, frames and ray counts still equal the oracle's — on awkward geometry, all sphere walks."""
    text = cases.random_world(seed, n_spheres=66, n_triangles=40)
    cam, world = ob.parse_input(text)
    W, H, spp, depth = 36, 24, 2, 6
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    cf = cam.floats()
    for flags in (0, 0x80000000, 0x20000000):
        out = np.zeros((H, W, 4), np.uint8)
        n = C.c_uint64()
        rc = hostsim_perturbed.hostsim_render(text.encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), W, H, spp, depth,
                                              ob.SEED_DEFAULT, flags, 0, 0, out.ctypes.data, C.byref(n))
        assert rc == 0 and n.value == rays, (seed, flags)
        assert np.array_equal(out, want), (seed, flags)


def test_c5_triangles_with_perturbed_reciprocals(hostsim_perturbed, ob, scenes):
    text = scenes.synthetic_world(200, 600, seed=10000)          # triangle-heavy cut of the C5 generator
    cam, world = ob.parse_input(text)
    W, H, spp, depth = 48, 27, 2, 16
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    out = np.zeros((H, W, 4), np.uint8)
    n = C.c_uint64()
    cf = cam.floats()
    rc = hostsim_perturbed.hostsim_render(text.encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), W, H, spp, depth,
                                          ob.SEED_DEFAULT, 0x80000000, 0, 0, out.ctypes.data, C.byref(n))
    assert rc == 0 and n.value == rays and np.array_equal(out, want)


def test_constant_divisors_of_the_slot_decode_equal_the_divide(hostsim):
    """rt_types.h rt_divisor / rt_div (the render kernel's slot decode divides by two launch constants): equal to `/` for
    every sub-tile count a frame of up to 65,536 x 65,536 pixels and any tile height can produce (a sample of them), all
    powers of two and their neighbours, numerators at every boundary and 20,000 random ones each."""
    import numpy as np
    widths = list(range(1, 700)) + [1920, 3840, 4096, 7680, 8192, 16384, 65535, 65536]
    ds = sorted({(w + 7) // 8 for w in widths} | {((w + 7) // 8) * r for w in widths for r in (1, 4, 16, 64, 4096)}
                | {3, 7, 641, 6700417, 2**31 - 1, 2**31})
    ds = [d for d in ds if 1 <= d <= 2**31]
    arr = (C.c_uint32 * len(ds))(*ds)
    assert hostsim.hostsim_divisor_mismatches(arr, len(ds), 20000, 12345) == 0


def _extreme_discriminant_world(variant: int) -> str:
    """Sphere groups whose discriminants leave [2^-100, 2^100] — the range the bare square-root sequences of
    sphere_group are used on: radii of 1e16..1e18 (disc ~ 1e32..1e36), radii whose square is denormal or zero in f32
    (disc a rounding residue: tiny, zero or negative), next to ordinary spheres of the same group, some of them
    concentric / coincident so that equal roots meet the reference's list-order tie rule."""
    big = ["10000000000000000.0", "300000000000000000.0", "1000000000000000000.0"][variant % 3]
    tiny = ["0.00000000000000000001", "0.000000000000000000000001", "0.0000000000000000000000000001"][variant % 3]
    lines = ["camera origin 0.0 0.25 1.0 aspect 1.5;",
             "material G : Diffuse color 0.5 0.6 0.4;", "material M : Metal color 0.8 0.8 0.9 fuzz 0.1;",
             "material D : Dielectric ir 1.5;", "material R : Diffuse color 0.9 0.2 0.2;"]
    def sphere(x, y, z, r, m): lines.append(f"sphere center {x} {y} {z} radius {r} material {m};")
    sphere("0.0", "-" + big, "-1.0", big, "G")                 # a "ground" whose discriminant overflows the range
    sphere("0.0", "0.3", "-2.0", "0.5", "R")
    sphere("0.0", "0.25", "-1.5", tiny, "M")                  # r*r underflows: disc is a rounding residue
    sphere("1.0", "0.3", "-2.5", "0.5", "M")
    sphere("1.0", "0.3", "-2.5", "0.5", "D")                  # coincident with the previous one: equal roots, first wins
    sphere("-1.2", "0.4", "-2.2", "0.6", "D")
    sphere("-1.2", "0.4", "-2.2", "0.3", "R")                 # concentric, inside a dielectric
    sphere("0.0", big, "-1.0", big, "G")                       # a "ceiling" of the same size: two huge discs in one group
    if variant >= 3:                                           # a second group: ordinary spheres + one more tiny one
        for i in range(7):
            sphere(f"{-2.0 + 0.6 * i:.1f}", "0.1", "-3.5", "0.25", "RMD"[i % 3])
        sphere("0.5", "0.25", "-1.2", tiny, "R")
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("variant", range(6))
def test_sphere_groups_with_out_of_range_discriminants(hostsim, ob, variant):
    """rt_trace.cuh sphere_group: the pair-wise root finding serves groups whose discriminants all lie in the proven range;
    any other group takes the per-sphere path and its sqrtf tail with the tie rule spelled out.  Both must give the
    oracle's frame and ray count on groups that mix ordinary, overflowing and vanishing discriminants."""
    text = _extreme_discriminant_world(variant)
    cam, world = ob.parse_input(text)
    W, H, spp, depth = 48, 32, 3, 8
    want, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth)
    assert rays > W * H * spp                                   # something is hit and bounces
    cf = cam.floats()
    out = np.zeros((H, W, 4), np.uint8)
    n = C.c_uint64()
    rc = hostsim.hostsim_render(text.encode(), cf.ctypes.data_as(C.POINTER(C.c_float)), W, H, spp, depth,
                                ob.SEED_DEFAULT, 0, 0, 0, out.ctypes.data, C.byref(n))
    assert rc == 0 and n.value == rays
    assert np.array_equal(out, want)


def test_work_space_decode_is_a_bijection_onto_the_frame(hostsim):
    """rt_trace.cuh decode_slot (8x4 sub-tiles, bottom-up strips, boustrophedon tile deal, pass-major fused passes, sample
    items, compact shard buffers, the constant-divisor arithmetic): over all shards every (pixel, pass, sample) is produced
    exactly once, by the shard that owns its tile, with the right output index — for ragged widths and heights, every
    tile height up to 64, 1-8 shards."""
    import itertools
    sizes = [(1, 1), (7, 5), (8, 4), (9, 17), (33, 31), (64, 64), (100, 37), (257, 66), (401, 225), (1921, 130)]
    n = 0
    for (W, H), tile_rows, shards in itertools.product(sizes, (4, 8, 16, 20, 64), (1, 2, 3, 8)):
        for passes, spp, items, compact in ((1, 1, 0, 0), (3, 2, 0, 0), (1, 3, 1, 0), (1, 1, 0, 1), (2, 1, 0, 1)):
            if W * H * passes * max(spp if items else 1, 1) > 600_000:
                continue
            assert hostsim.hostsim_decode_violations(W, H, tile_rows, shards, passes, spp, items, compact) == 0, \
                (W, H, tile_rows, shards, passes, spp, items, compact)
            n += 1
    assert n > 500
