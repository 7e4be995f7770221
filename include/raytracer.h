/*
 * raytracer.h — drop-in C ABI of the raytracer crate, served by the B200-native library
 * (libraytracer.so built from rust-swift-raytracer_b200/csrc/).
 *
 * Every declaration below is binary- and source-compatible with the header cbindgen 0.18
 * generates for the reference crate (type prefix `Rust_`, functions un-prefixed):
 *   reference header : MacOSPlatform/MacOSPlatform/Engine/includes/raytracer.h:1-47
 *   generated from   : raytracer/src/lib.rs:22-63, raytracer/src/color.rs:3-10,
 *                      raytracer/src/maths.rs:53-55,98-103, raytracer/cbindgen.toml:9,61
 * A caller compiled against the reference header (the Swift bridging header, a C program,
 * examples/c_raytracer.rs through `extern "C"`) links against this library unchanged.
 *
 * Additive entry points (options, stats, error reporting, frees) live in raytracer_b200.h.
 */
#ifndef RAYTRACER_H
#define RAYTRACER_H

#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque.  Replaces camera.rs:8-15 `Camera` behind Box<Camera> (lib.rs:32). */
typedef struct Rust_Camera Rust_Camera;

/* Opaque.  Replaces common.rs:227-230 `World` behind Box<World> (lib.rs:31); additionally
 * owns the packed device copies of the scene. */
typedef struct Rust_World Rust_World;

/* lib.rs:29-33 (#[repr(C)], two Box pointers).  C callers read both fields and overwrite
 * `camera` with the result of move_camera_position (GameView.swift:200-216). */
typedef struct Rust_WorldHandle {
  struct Rust_World *world;
  struct Rust_Camera *camera;
} Rust_WorldHandle;

/* color.rs:3-10 (#[repr(C)]): bytes R,G,B,A in memory. */
typedef struct Rust_ColorU8 {
  uint8_t r;
  uint8_t g;
  uint8_t b;
  uint8_t a;
} Rust_ColorU8;

/* lib.rs:22-27 (#[repr(C)]): row-major, top row first, width*height pixels. */
typedef struct Rust_CFramebuffer {
  size_t width;
  size_t height;
  struct Rust_ColorU8 *pixels;
} Rust_CFramebuffer;

/* maths.rs:98-103 (#[repr(C)]) */
typedef struct Rust_NVec3 {
  float x;
  float y;
  float z;
} Rust_NVec3;

#ifndef __cplusplus
/* maths.rs:53-55 */
#define Rust_X_AXIS (Rust_NVec3){ .x = 1.0, .y = 0.0, .z = 0.0 }
#define Rust_Y_AXIS (Rust_NVec3){ .x = 0.0, .y = 1.0, .z = 0.0 }
#define Rust_Z_AXIS (Rust_NVec3){ .x = 0.0, .y = 0.0, .z = 1.0 }
#else
#define Rust_X_AXIS (Rust_NVec3{ 1.0f, 0.0f, 0.0f })
#define Rust_Y_AXIS (Rust_NVec3{ 0.0f, 1.0f, 0.0f })
#define Rust_Z_AXIS (Rust_NVec3{ 0.0f, 0.0f, 1.0f })
#endif

/* Replaces lib.rs:37-46.  `source`: NUL-terminated UTF-8 world text (grammar
 * parser.rs:326-335).  Parses it, packs the primitives SoA and returns a heap handle owned
 * by the caller (free with rt_free_world; the reference exports no destructor and leaks).
 * Where the reference panics (`.unwrap()` x2 on invalid UTF-8 / any ParseError, lib.rs:40)
 * this returns NULL and rt_last_error() names the ParseError. */
struct Rust_WorldHandle *load_world(const char *source);

/* Replaces lib.rs:60-63.  Consumes (frees) `camera` and returns a new heap camera
 * Camera::new_at(old.position + (x,y,z), old.aspect_ratio()). */
struct Rust_Camera *move_camera_position(struct Rust_Camera *camera, float x, float y, float z);

/* Replaces lib.rs:49-57: one frame at 16 samples per pixel, max depth 8 (lib.rs:51),
 * rendered by the CUDA path on the current device.  The frame is written into
 * framebuffer.pixels (the caller's width*height allocation, which every reference caller
 * provides: GameView.swift:125-129,350-354, c_raytracer.rs:53-58) and the same struct is
 * returned — a defined superset of the reference, whose returned `pixels` dangles
 * (lib.rs:79-87).  Reads handle->camera at call time.  On failure (no CUDA device, CUDA
 * error) the pixels are left untouched and rt_last_error() is set.
 * Environment, for callers that cannot pass options: RT_GPUS=N renders the frame on N GPUs
 * (row tiles, stored straight into device 0's frame over NVLink); RT_DETERMINISTIC=1 uses the
 * fixed sub-pixel offset (0.5, 0.5) instead of the two jitter draws. */
struct Rust_CFramebuffer render(struct Rust_CFramebuffer framebuffer,
                                const struct Rust_WorldHandle *handle);

#ifdef __cplusplus
}
#endif
#endif /* RAYTRACER_H */
