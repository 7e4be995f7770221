"""ctypes binding of the CPU oracle (oracle/rt_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package
(rust-swift-raytracer_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "build" / "librt_oracle.so"

SEED_DEFAULT = 2547549  # random.rs:9

DIFFUSE, METAL, DIELECTRIC, EMISSION = 0, 1, 2, 3
RNG_SERIAL, RNG_PER_SAMPLE = 0, 1


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def tuple(self):
        return (self.x, self.y, self.z)


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("r", C.c_float), ("g", C.c_float), ("b", C.c_float),
                ("param", C.c_float)]


class Sphere(C.Structure):
    _fields_ = [("center", Vec3), ("radius", C.c_float), ("material", Material)]


class Triangle(C.Structure):
    _fields_ = [("v0", Vec3), ("v1", Vec3), ("v2", Vec3), ("normal", Vec3), ("material", Material)]


class Camera(C.Structure):
    _fields_ = [("origin", Vec3), ("lower_left_corner", Vec3), ("horizontal", Vec3),
                ("vertical", Vec3)]

    def floats(self):
        return np.frombuffer(bytes(self), dtype=np.float32).copy()


class Options(C.Structure):
    _fields_ = [("samples_per_pixel", C.c_int32), ("max_ray_bounces", C.c_int32),
                ("rng_mode", C.c_int32), ("seed", C.c_uint32), ("fixed_jitter", C.c_int32),
                ("sample_begin", C.c_int32), ("threads", C.c_int32), ("reserved", C.c_int32)]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (building the checker is not using it)."""
    src_m = max((_HERE / n).stat().st_mtime for n in ("rt_oracle.c", "rt_oracle.h", "unicode_alnum.h", "Makefile"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_LIB_PATH))
    u32p = C.POINTER(C.c_uint32)
    L.orc_xorshift32.restype = C.c_uint32
    L.orc_xorshift32.argtypes = [u32p]
    L.orc_random_f32.restype = C.c_float
    L.orc_random_f32.argtypes = [u32p]
    L.orc_random_bilateral_f32.restype = C.c_float
    L.orc_random_bilateral_f32.argtypes = [u32p]
    L.orc_sample_seed.restype = C.c_uint32
    L.orc_sample_seed.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    for name, n in (("orc_negate", 1), ("orc_normalize", 1), ("orc_cross", 2), ("orc_reflect", 2),
                    ("orc_project", 2)):
        f = getattr(L, name)
        f.restype = Vec3
        f.argtypes = [Vec3] * n
    L.orc_refract.restype = Vec3
    L.orc_refract.argtypes = [Vec3, Vec3, C.c_float]
    L.orc_f32_as_u8.restype = C.c_uint8
    L.orc_f32_as_u8.argtypes = [C.c_float]

    L.orc_camera_new_at.restype = Camera
    L.orc_camera_new_at.argtypes = [Vec3, C.c_float]
    L.orc_camera_new_with_vertical_fov.restype = Camera
    L.orc_camera_new_with_vertical_fov.argtypes = [Vec3, C.c_float, C.c_float]
    L.orc_camera_new_look_at.restype = C.c_int
    L.orc_camera_new_look_at.argtypes = [Vec3, Vec3, Vec3, C.c_float, C.c_float, C.POINTER(Camera)]
    L.orc_camera_aspect_ratio.restype = C.c_float
    L.orc_camera_aspect_ratio.argtypes = [C.POINTER(Camera)]
    L.orc_move_camera_position.restype = Camera
    L.orc_move_camera_position.argtypes = [C.POINTER(Camera), C.c_float, C.c_float, C.c_float]
    L.orc_cast_ray.restype = None
    L.orc_cast_ray.argtypes = [C.POINTER(Camera), C.c_float, C.c_float, C.POINTER(Vec3),
                               C.POINTER(Vec3)]

    L.orc_world_new.restype = C.c_void_p
    L.orc_world_free.argtypes = [C.c_void_p]
    L.orc_world_add_sphere.argtypes = [C.c_void_p, Vec3, C.c_float, Material]
    L.orc_world_add_triangle.argtypes = [C.c_void_p, Vec3, Vec3, Vec3, Material]
    L.orc_world_sphere_count.restype = C.c_size_t
    L.orc_world_sphere_count.argtypes = [C.c_void_p]
    L.orc_world_triangle_count.restype = C.c_size_t
    L.orc_world_triangle_count.argtypes = [C.c_void_p]
    L.orc_world_spheres.restype = C.POINTER(Sphere)
    L.orc_world_spheres.argtypes = [C.c_void_p]
    L.orc_world_triangles.restype = C.POINTER(Triangle)
    L.orc_world_triangles.argtypes = [C.c_void_p]
    L.orc_parse_input.restype = C.c_void_p
    L.orc_parse_input.argtypes = [C.c_char_p, C.POINTER(Camera), C.POINTER(C.c_int)]
    L.orc_parse_error_name.restype = C.c_char_p
    L.orc_parse_error_name.argtypes = [C.c_int]
    L.orc_world_hit.restype = C.c_int
    L.orc_world_hit.argtypes = [C.c_void_p, Vec3, Vec3, C.POINTER(C.c_float), C.POINTER(Vec3),
                                C.POINTER(Vec3), C.POINTER(C.c_int64)]
    L.orc_ray_trace.restype = C.c_int
    L.orc_ray_trace.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_void_p, C.c_size_t, C.c_size_t,
                                C.POINTER(Options), C.c_int32, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_uint64)]
    L.orc_ray_trace_rows.restype = C.c_int
    L.orc_ray_trace_rows.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_void_p, C.c_size_t, C.c_size_t,
                                     C.c_size_t, C.c_size_t, C.POINTER(Options), C.c_int32, C.c_void_p,
                                     C.c_void_p, C.POINTER(C.c_uint64)]
    L.orc_write_image.restype = C.c_int
    L.orc_write_image.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_char_p]
    _lib = L
    return L


def v3(x, y, z) -> Vec3:
    return Vec3(float(x), float(y), float(z))


def material(kind: int, color=(1.0, 1.0, 1.0), param: float = 0.0) -> Material:
    return Material(kind, float(color[0]), float(color[1]), float(color[2]), float(param))


class ParseError(Exception):
    pass


class World:
    """Owner of an orc_world*."""

    def __init__(self, ptr=None):
        self._L = lib()
        self.ptr = ptr if ptr is not None else self._L.orc_world_new()

    def __del__(self):
        if getattr(self, "ptr", None):
            self._L.orc_world_free(self.ptr)
            self.ptr = None

    def add_sphere(self, center, radius, mat: Material):
        self._L.orc_world_add_sphere(self.ptr, v3(*center), float(radius), mat)

    def add_triangle(self, p0, p1, p2, mat: Material):
        self._L.orc_world_add_triangle(self.ptr, v3(*p0), v3(*p1), v3(*p2), mat)

    @property
    def n_spheres(self):
        return self._L.orc_world_sphere_count(self.ptr)

    @property
    def n_triangles(self):
        return self._L.orc_world_triangle_count(self.ptr)

    def spheres(self):
        p = self._L.orc_world_spheres(self.ptr)
        return [p[i] for i in range(self.n_spheres)]

    def triangles(self):
        p = self._L.orc_world_triangles(self.ptr)
        return [p[i] for i in range(self.n_triangles)]

    def hit(self, origin, direction):
        t = C.c_float()
        pos, nrm = Vec3(), Vec3()
        prim = C.c_int64(-1)
        ok = self._L.orc_world_hit(self.ptr, v3(*origin), v3(*direction), C.byref(t), C.byref(pos),
                                   C.byref(nrm), C.byref(prim))
        if not ok:
            return None
        return t.value, pos.tuple(), nrm.tuple(), prim.value


def parse_input(source: str | bytes):
    """parser.rs:336-382 -> (Camera, World); raises ParseError."""
    L = lib()
    if isinstance(source, str):
        source = source.encode("utf-8")
    cam = Camera()
    err = C.c_int(0)
    ptr = L.orc_parse_input(source, C.byref(cam), C.byref(err))
    if not ptr:
        raise ParseError(L.orc_parse_error_name(err.value).decode())
    return cam, World(ptr)


def camera_new_at(origin, aspect) -> Camera:
    return lib().orc_camera_new_at(v3(*origin), float(aspect))


def camera_new_with_vertical_fov(origin, vfov, aspect) -> Camera:
    return lib().orc_camera_new_with_vertical_fov(v3(*origin), float(vfov), float(aspect))


def camera_new_look_at(origin, look_at, up, vfov, aspect) -> Camera:
    cam = Camera()
    rc = lib().orc_camera_new_look_at(v3(*origin), v3(*look_at), v3(*up), float(vfov),
                                      float(aspect), C.byref(cam))
    if rc:
        raise ValueError("new_look_at assertion %d (camera.rs:50/:62)" % rc)
    return cam


def move_camera_position(cam: Camera, x, y, z) -> Camera:
    return lib().orc_move_camera_position(C.byref(cam), float(x), float(y), float(z))


def ray_trace(world: World, camera: Camera, width: int, height: int, spp: int, depth: int, *,
              rng_mode: int = RNG_PER_SAMPLE, seed: int = SEED_DEFAULT, fixed_jitter: bool = False,
              sample_begin: int = 0, threads: int | None = None, resolve_spp: int | None = None,
              accum_in: np.ndarray | None = None, want_accum: bool = False,
              rows: tuple[int, int] | None = None):
    """common.rs:320-361.  Returns (pixels[H,W,4] uint8, ray_count, accum or None).
    rows=(begin, end): only image rows [begin, end) (0 = top) are traced (per-sample RNG mode); the
    rest of the returned frame stays zero and ray_count is the band's."""
    L = lib()
    if threads is None:
        threads = os.cpu_count() or 1
    opt = Options(int(spp), int(depth), int(rng_mode), int(seed) & 0xFFFFFFFF, int(bool(fixed_jitter)),
                  int(sample_begin), int(threads), 0)
    pixels = np.zeros((height, width, 4), dtype=np.uint8)
    accum = np.zeros((height, width, 4), dtype=np.float32) if want_accum else None
    if accum_in is not None:
        accum_in = np.ascontiguousarray(accum_in, dtype=np.float32)
        assert accum_in.shape == (height, width, 4)
    rays = C.c_uint64(0)
    tail = (C.byref(opt), int(resolve_spp if resolve_spp is not None else spp),
            accum_in.ctypes.data if accum_in is not None else None,
            accum.ctypes.data if accum is not None else None, C.byref(rays))
    if rows is not None:
        assert rng_mode == RNG_PER_SAMPLE, "a row band is only defined for the per-sample RNG mode"
        rc = L.orc_ray_trace_rows(world.ptr, C.byref(camera), pixels.ctypes.data, width, height,
                                  int(rows[0]), int(rows[1]), *tail)
    else:
        rc = L.orc_ray_trace(world.ptr, C.byref(camera), pixels.ctypes.data, width, height, *tail)
    if rc:
        raise RuntimeError("orc_ray_trace failed: %d" % rc)
    return pixels, rays.value, accum


def write_image(pixels: np.ndarray, path: str):
    h, w, _ = pixels.shape
    px = np.ascontiguousarray(pixels)
    if lib().orc_write_image(px.ctypes.data, w, h, path.encode()):
        raise OSError("orc_write_image failed")
