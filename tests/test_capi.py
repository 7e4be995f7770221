"""The C-ABI library: loads without a GPU, exports every symbol include/*.h declares, keeps the
reference's struct layouts, and fails loudly (never falls back) when no CUDA device exists."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions():
    names = set()
    for header in ("raytracer.h", "raytracer_b200.h"):
        text = (ROOT / "include" / header).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"#define[^\n]*\n", "\n", text)
        for m in re.finditer(r"\b([a-z_][a-z0-9_]*)\s*\(", text):
            names.add(m.group(1))
    return names - {"sizeof", "defined"}


def test_library_exports_every_declared_symbol(rt):
    L = rt.lib()
    declared = _declared_functions()
    assert {"load_world", "render", "move_camera_position"} <= declared
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/*.h but not exported"
    assert declared == set(rt.EXPORTED_SYMBOLS)


def test_reference_struct_layouts(rt):
    # raytracer.h:12-28 of the reference (cbindgen output)
    assert C.sizeof(rt._ColorU8) == 4
    assert C.sizeof(rt._CFramebuffer) == 24 and rt._CFramebuffer.pixels.offset == 16
    assert C.sizeof(rt._WorldHandle) == 16 and rt._WorldHandle.camera.offset == 8
    assert rt.lib().rt_abi_version() == 2
    # the Python mirrors of the additive structs are the header's layouts (ABI version 2)
    assert C.sizeof(rt.RenderStats) == 80 and C.sizeof(rt._RenderOptions) == 72 and C.sizeof(rt.PeerQueue) == 16


def test_reference_header_is_source_compatible(tmp_path):
    """A C caller written against the reference header compiles against include/raytracer.h."""
    import subprocess
    src = tmp_path / "caller.c"
    src.write_text('#include "raytracer.h"\n'
                   "int main(void) {\n"
                   "  Rust_ColorU8 px[4]; Rust_CFramebuffer fb = { 2, 2, px };\n"
                   "  Rust_WorldHandle *h = load_world(\"camera origin 0.0 0.0 0.0 aspect 1.0;\");\n"
                   "  if (!h) return 1;\n"
                   "  h->camera = move_camera_position(h->camera, 1.0f, 0.0f, 0.0f);\n"
                   "  Rust_NVec3 y = Rust_Y_AXIS; (void)y;\n"
                   "  (void)fb; (void)render; return 0; }\n")
    lib = ROOT / "rust-swift-raytracer_b200" / "lib"
    exe = tmp_path / "caller"
    r = subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), f"-L{lib}",
                        "-lraytracer", f"-Wl,-rpath,{lib}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_additive_header_is_strict_c_and_cxx(tmp_path):
    """include/raytracer_b200.h (plain pointers and sizes only) compiles as pedantic C11 and as C++17."""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('#include "raytracer_b200.h"\n'
                   "int main(void) {\n"
                   "  RtRenderOptions o; RtRenderStats s; (void)s;\n"
                   "  o.struct_size = (uint32_t)sizeof o; o.flags = RT_OPT_FIXED_JITTER | RT_OPT_GROUP_CULL;\n"
                   "  return rt_abi_version() == RT_B200_ABI_VERSION && o.flags ? 0 : 1; }\n")
    lib = ROOT / "rust-swift-raytracer_b200" / "lib"
    for cc, std in (("/usr/bin/gcc", ["-std=c11", "-pedantic"]), ("/usr/bin/g++", ["-std=c++17", "-x", "c++"])):
        exe = tmp_path / ("use_" + Path(cc).name.replace("+", "p"))
        r = subprocess.run([cc] + std + ["-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(src), f"-L{lib}",
                            "-lraytracer", f"-Wl,-rpath,{lib}", "-o", str(exe)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert subprocess.run([str(exe)]).returncode == 0


def test_move_camera_position_lib_rs_60(rt, ob, scenes):
    h = rt.load_world(scenes.default_world())
    h.set_camera_look_at((1, 2, 3), (0, 0, -1), (0, 1, 0), 0.8, 1.5)
    cam = ob.camera_new_look_at((1, 2, 3), (0, 0, -1), (0, 1, 0), 0.8, 1.5)
    assert np.array_equal(h.camera_floats(), cam.floats())
    rt.move_camera_position(h, 0.5, -1.0, 2.0)
    assert np.array_equal(h.camera_floats(), ob.move_camera_position(cam, 0.5, -1.0, 2.0).floats())


def test_camera_constructors_match_oracle(rt, ob, scenes):
    h = rt.load_world(scenes.default_world())
    h.set_camera_at((0.25, -1.0, 4.0), 1.3333)
    assert np.array_equal(h.camera_floats(), ob.camera_new_at((0.25, -1.0, 4.0), 1.3333).floats())
    h.set_camera_vertical_fov((0.25, -1.0, 4.0), 1.1, 1.3333)
    assert np.array_equal(h.camera_floats(), ob.camera_new_with_vertical_fov((0.25, -1.0, 4.0), 1.1, 1.3333).floats())
    with pytest.raises(rt.RenderError):
        h.set_camera_look_at((0, 0, 0), (0, 0, 0), (0, 1, 0), 1.0, 1.0)     # camera.rs:50
    with pytest.raises(rt.RenderError):
        h.set_camera_look_at((0, 0, 0), (0, 1, 0), (0, 1, 0), 1.0, 1.0)     # camera.rs:62


def test_set_camera_raw_round_trips(rt, scenes):
    h = rt.load_world(scenes.default_world())
    cam = np.arange(12, dtype=np.float32) * np.float32(0.37) - np.float32(1.5)
    h.set_camera_raw(cam)
    assert np.array_equal(h.camera_floats(), cam)


def test_world_builder_reaches_emission(rt):
    h = rt.world_new((0, 0, 0), 1.5)
    h.add_sphere((0, 0, -1), 0.5, rt.EMISSION, (2.0, 1.0, 0.5))
    h.add_triangle((-1, -1, -2), (1, -1, -2), (0, 1, -2), rt.METAL, (0.5, 0.5, 0.5), 0.1)
    assert (h.n_spheres, h.n_triangles) == (1, 1)
    assert h.sphere(0)[4] == rt.EMISSION
    assert np.array_equal(h.triangle(0)[9:12], np.array([0, 0, 1], np.float32))


def test_no_cpu_fallback(rt, scenes):
    """Without a CUDA device every render entry point must fail loudly."""
    if rt.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = rt.load_world(scenes.default_world())
    fb = rt.Framebuffer(8, 6)
    fb.pixels[...] = 7
    with pytest.raises(rt.RenderError, match="no CUDA device"):
        rt.render(fb, h)
    assert (fb.pixels == 7).all()          # untouched on failure
    with pytest.raises(rt.RenderError):
        rt.render_with_options(fb, h, rt.Options(1, 1))
    with pytest.raises(rt.RenderError):
        rt.measure_fp32_peak()


def test_product_does_not_import_the_oracle():
    pkg = ROOT / "rust-swift-raytracer_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + \
            list(pkg.rglob("*.h")) + list(pkg.rglob("*.hpp")):
        text = f.read_text()
        assert "oracle_binding" not in text and "rt_oracle" not in text and "hostsim" not in text.replace(
            "tests/hostsim", ""), f


def test_write_image_p3(rt, tmp_path):
    fb = rt.Framebuffer(3, 2)
    fb.pixels[...] = np.arange(24, dtype=np.uint8).reshape(2, 3, 4)
    rt.write_image(fb, tmp_path / "a.ppm")
    lines = (tmp_path / "a.ppm").read_text().split("\n")
    assert lines[:3] == ["P3", "3 2", "255"]                       # image.rs:68-71
    assert lines[3] == "0 1 2" and lines[8] == "20 21 22" and lines[9] == ""
    rt.write_image(fb, tmp_path / "b.ppm", binary=True)
    raw = (tmp_path / "b.ppm").read_bytes()
    assert raw.startswith(b"P6\n3 2\n255\n") and raw[-18:] == fb.pixels[:, :, :3].tobytes()


def test_write_image_equals_the_oracle_writer_byte_for_byte(rt, ob, tmp_path):
    """image.rs:59-81: the table-driven P3 writer produces exactly the bytes of the oracle's fprintf restatement —
    every channel value 0..255 in every position, sizes that straddle the 1 MiB output buffer — and the binary P6
    file carries the same pixels."""
    rng = np.random.default_rng(7)
    for W, H in ((16, 16), (257, 3), (701, 523)):
        fb = rt.Framebuffer(W, H)
        fb.pixels[...] = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        if (W, H) == (16, 16):
            fb.pixels[:, :, 0] = np.arange(256, dtype=np.uint8).reshape(16, 16)
            fb.pixels[:, :, 1] = np.arange(256, dtype=np.uint8).reshape(16, 16)[::-1]
            fb.pixels[:, :, 2] = np.arange(256, dtype=np.uint8).reshape(16, 16).T
        rt.write_image(fb, tmp_path / "got.ppm")
        ob.write_image(fb.pixels, str(tmp_path / "want.ppm"))
        assert (tmp_path / "got.ppm").read_bytes() == (tmp_path / "want.ppm").read_bytes()
        rt.write_image(fb, tmp_path / "got6.ppm", binary=True)
        raw = (tmp_path / "got6.ppm").read_bytes()
        head = b"P6\n%d %d\n255\n" % (W, H)
        assert raw.startswith(head) and raw[len(head):] == fb.pixels[:, :, :3].tobytes()


def test_invalid_options_fail_before_any_device_work_and_leave_the_frame_alone(rt, scenes):
    """Errors never cross the ABI: a bad option block makes render_with_options fail with a text in
    rt_last_error() — decided on the host, so this holds on a box without a GPU — and the caller's pixels are untouched."""
    import numpy as np
    h = rt.load_world(scenes.default_world())
    fb = rt.Framebuffer(24, 16)
    fb.pixels[...] = 0xA5
    cases_ = [
        (dict(tile_rows=6), "tile_rows"), (dict(tile_rows=2), "tile_rows"),       # (0 means the default, 16)
        (dict(shard_index=2, shard_count=2), "shard"),
        (dict(passes=3), "multiple of passes"),
        (dict(accum_in=True), "accum"),
        (dict(peer_queues=[(4096, i) for i in range(9)], shard_count=9, full_frame_out=True), "too many peer queues"),
        (dict(peer_queues=[(4096, 0), (8192, 1)], shard_count=3, full_frame_out=True), "every shard"),
        (dict(peer_queues=[(4096, 0), (8192, 1)], shard_count=2), "full-frame"),
    ]
    for kw, needle in cases_:
        with pytest.raises(rt.RenderError) as e:
            rt.render_with_options(fb, h, rt.Options(4, 4, **kw))
        assert needle in str(e.value), (kw, str(e.value))
        assert needle in rt.last_error()
        assert np.all(fb.pixels == 0xA5), kw


def test_load_world_reports_the_parse_error_instead_of_panicking(rt):
    """lib.rs:37-46 unwraps the parse result (a panic across the FFI); here load_world returns NULL and
    rt_last_error() names the reference's error class."""
    for text in ("sphere center 0 0 0 radius 1 material nope;", "camera origin 0 0;", "material M : Shiny;", "\xff"):
        with pytest.raises(rt.ParseError) as e:
            rt.load_world(text)
        assert "ParseError" in str(e.value) and rt.last_error() == str(e.value)
