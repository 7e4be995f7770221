#!/usr/bin/env python3
"""Put the B200 render path behind the UNMODIFIED `raytracer` crate (Rust callers).

    python bindings/rust/apply_shim.py /path/to/Rust-Swift-Raytracer/raytracer
    RAYTRACER_B200_LIB_DIR=/path/to/rust-swift-raytracer_b200/lib cargo run --release      # src/main.rs, unchanged
    RAYTRACER_B200_LIB_DIR=...                                  cargo run --example c_raytracer  # from the workspace root

What it edits (anchors are the reference's own signatures; nothing else is touched):
  build.rs        + link search path / rpath for libraytracer.so               (build.rs:10)
  src/ffi_b200.rs   new file (copied from this directory): the extern "C" block and owned handle
  src/lib.rs        load_world / render / move_camera_position (lib.rs:37-63) stop being #[no_mangle] definitions —
                    libraytracer.so defines those symbols now — and become thin Rust wrappers of the same names
  src/camera.rs   + Camera::b200_floats(): the 12 floats of camera.rs:8-15 (fields are private to the module)
  src/common.rs     World::new (common.rs:233-235) mirrors its primitives into the library once, in list order;
                    ray_trace (common.rs:320-361) keeps its signature and forwards to render_with_options

The authoring environment has no rustc/cargo: this script and ffi_b200.rs are NOT compiled there.  The C side of
every call they make is covered by the repository's tests (tests/c_caller, tests/test_capi.py, test_gpu_parity.py).
"""
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent


def edit(path: Path, fn):
    text = path.read_text()
    new = fn(text)
    if new == text:
        raise SystemExit(f"{path}: anchor not found — is this the unmodified reference crate?")
    path.write_text(new)
    print("patched", path)


def replace_once(text, old, new):
    if text.count(old) != 1:
        raise SystemExit(f"anchor occurs {text.count(old)} times: {old[:60]!r}")
    return text.replace(old, new)


def build_rs(s):
    return replace_once(s, 'fn main() {\n', '''fn main() {
    // B200 render path: link libraytracer.so (built by `python rust-swift-raytracer_b200/build.py`)
    if let Ok(dir) = env::var("RAYTRACER_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rerun-if-env-changed=RAYTRACER_B200_LIB_DIR");

''')


def camera_rs(s):
    anchor = '    pub fn position(&self) -> Vec3 {\n        self.origin\n    }\n'
    return replace_once(s, anchor, anchor + '''
    /// origin, lower_left_corner, horizontal, vertical — what the B200 library's rt_set_camera_raw installs.
    pub(crate) fn b200_floats(&self) -> [f32; 12] {
        [self.origin.x, self.origin.y, self.origin.z,
         self.lower_left_corner.x, self.lower_left_corner.y, self.lower_left_corner.z,
         self.horizontal.x, self.horizontal.y, self.horizontal.z,
         self.vertical.x, self.vertical.y, self.vertical.z]
    }
''')


LIB_WRAPPERS = '''// The three C exports (`load_world`, `render`, `move_camera_position`) are now DEFINED by libraytracer.so
// (the B200 render path); the crate keeps Rust functions of the same names and meaning for its Rust callers
// (examples/c_raytracer.rs) and no longer exports #[no_mangle] symbols of its own.
pub use ffi_b200::{OwnedWorldHandle, WorldHandle};

pub fn load_world(source: *const c_char) -> OwnedWorldHandle {
    let handle = unsafe { ffi_b200::load_world(source) };
    match NonNull::new(handle) {
        Some(h) => OwnedWorldHandle(h),
        None => panic!("load_world: {}", ffi_b200::last_error()),     // the reference unwrap()s here too
    }
}

pub fn render(framebuffer: CFramebuffer, handle: *const WorldHandle) -> CFramebuffer {
    let fb = ffi_b200::CFramebuffer { width: framebuffer.width, height: framebuffer.height, pixels: framebuffer.pixels };
    let out = unsafe { ffi_b200::render(fb, handle) };               // 16 spp, depth 8; frame lands in `pixels`
    CFramebuffer { width: out.width, height: out.height, pixels: out.pixels }
}

pub fn move_camera_position(handle: &mut WorldHandle, x: f32, y: f32, z: f32) {
    handle.camera = unsafe { ffi_b200::move_camera_position(handle.camera, x, y, z) };
}

'''


def lib_rs(s):
    s = replace_once(s, 'pub mod color;\n', 'pub mod color;\npub mod ffi_b200;\n')
    # the reference's own WorldHandle (two Boxes) gives way to the library's (two opaque pointers)
    a = s.index('#[repr(C)]\npub struct WorldHandle {')
    b = s.index('impl Into<Framebuffer> for CFramebuffer {')
    s = s[:a] + LIB_WRAPPERS + s[b:]
    for unused in ('use camera::Camera;\n', 'use common::{World, Options, ray_trace};\n', 'use std::ffi::CStr;\n', 'use maths::Vec3;\n'):
        s = s.replace(unused, '')
    return s


WORLD_NEW = '''fn b200_material(material: &MaterialType) -> (u32, [f32; 3], f32) {
    use crate::ffi_b200::*;
    match material {
        MaterialType::Diffuse(c)    => (RT_MATERIAL_DIFFUSE,    [c.r, c.g, c.b], 0.0),
        MaterialType::Metal(c, f)   => (RT_MATERIAL_METAL,      [c.r, c.g, c.b], *f),
        MaterialType::Dielectric(i) => (RT_MATERIAL_DIELECTRIC, [1.0, 1.0, 1.0], *i),
        MaterialType::Emission(c)   => (RT_MATERIAL_EMISSION,   [c.r, c.g, c.b], 0.0),
    }
}

impl World {
    pub fn new(spheres: Vec<Sphere>, meshes: Vec<Mesh>) -> Self {
        use crate::ffi_b200::*;
        let origin = [0.0f32; 3];
        let handle = unsafe { rt_world_new(origin.as_ptr(), 1.0) };
        let b200 = OwnedWorldHandle(std::ptr::NonNull::new(handle).expect("rt_world_new"));
        for s in &spheres {
            let (kind, color, param) = b200_material(&s.material);
            let c = [s.center.x, s.center.y, s.center.z];
            let rc = unsafe { rt_world_add_sphere(handle, c.as_ptr(), s.radius, kind, color.as_ptr(), param) };
            assert!(rc == 0, "rt_world_add_sphere: {}", last_error());
        }
        for mesh in &meshes {               // World::hit walks the meshes in order; the library holds one list
            for t in &mesh.triangles {
                let (kind, color, param) = b200_material(&t.material);
                let (v0, v1, v2) = ([t.v0.x, t.v0.y, t.v0.z], [t.v1.x, t.v1.y, t.v1.z], [t.v2.x, t.v2.y, t.v2.z]);
                let rc = unsafe { rt_world_add_triangle(handle, v0.as_ptr(), v1.as_ptr(), v2.as_ptr(), kind, color.as_ptr(), param) };
                assert!(rc == 0, "rt_world_add_triangle: {}", last_error());
            }
        }
        Self { spheres, meshes, b200 }
    }
'''

RAY_TRACE = '''/// The per-pixel render loop, on the GPU: forwards to libraytracer.so (`render_with_options`).  Same signature,
/// same frame layout; the image is the per-(pixel, sample)-seeded realisation of the same estimator.
pub fn ray_trace(world: &World, camera: &Camera, mut framebuffer: Framebuffer, options: &mut Options) -> Framebuffer {
    use crate::ffi_b200::*;
    let handle = world.b200.0.as_ptr();
    let cam = camera.b200_floats();
    let rc = unsafe { rt_set_camera_raw(handle, cam.as_ptr()) };
    assert!(rc == 0, "rt_set_camera_raw: {}", last_error());
    framebuffer.pixels.resize(framebuffer.width * framebuffer.height, ColorU8 { r: 0, g: 0, b: 0, a: 0 });
    let opt = RtRenderOptions::new(options.samples_per_pixel, options.max_ray_bounces);
    let fb = CFramebuffer {
        width: framebuffer.width,
        height: framebuffer.height,
        pixels: std::ptr::NonNull::new(framebuffer.pixels.as_mut_ptr()).unwrap(),
    };
    unsafe { render_with_options(fb, handle, &opt) };
    let err = last_error();
    assert!(err.is_empty(), "render_with_options: {}", err);
    if let Some(logger) = &mut options.logger {
        write!(logger, "\\rScanline: {:<4}", 0).unwrap();
    }
    framebuffer
}
'''


def common_rs(s):
    s = replace_once(s, '    meshes:  Vec<Mesh>,\n}\n', '    meshes:  Vec<Mesh>,\n    b200:    crate::ffi_b200::OwnedWorldHandle,'
                     '     // the same primitives, in list order, on the library side\n}\n')
    a = s.index('impl World {\n    pub fn new(')
    b = s.index('    pub fn hit(&self, ray: &Ray) -> Option<HitRecord> {')
    s = s[:a] + WORLD_NEW + '\n' + s[b:]
    a = s.index('pub fn ray_trace(world: &World')
    return s[:a] + RAY_TRACE


def main():
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    crate = Path(sys.argv[1]).resolve()
    if not (crate / "src" / "common.rs").exists():
        raise SystemExit(f"{crate} does not look like the raytracer crate")
    shutil.copy(HERE / "ffi_b200.rs", crate / "src" / "ffi_b200.rs")
    print("copied", crate / "src" / "ffi_b200.rs")
    edit(crate / "build.rs", build_rs)
    edit(crate / "src" / "camera.rs", camera_rs)
    edit(crate / "src" / "lib.rs", lib_rs)
    edit(crate / "src" / "common.rs", common_rs)


if __name__ == "__main__":
    main()
