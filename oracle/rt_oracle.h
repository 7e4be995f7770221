/*
 * rt_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY — never linked into, imported
 * by, or called from the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it).
 *
 * A plain-C, single-file restatement of the reference raytracer crate's per-pixel
 * render loop (raytracer/src/common.rs:320-361 and everything it calls).
 *
 * PARITY STATUS: "image-level parity unpinned".  The reference ships no golden
 * image, pixel hash or per-ray expected value, and it cannot be compiled in this
 * environment (no rustc/cargo; prebuilt libraytracer.a stripped).  What IS pinned:
 *   - the reference's own known-answer tests on this path: maths.rs:244-249
 *     (negate), :252-257 (reflect), :280-286 (refract)   -> tests/test_oracle_kat.py
 *   - the xorshift32 sequence derived from random.rs:8-30 by exact integer
 *     arithmetic                          -> tests/test_oracle_kat.py
 *   - an independent second restatement in numpy float32 (oracle/rt_oracle_np.py)
 *     that must agree with this file bit-for-bit          -> tests/test_oracle_cross.py
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z; } orc_vec3;

/* materials.rs:7-12 */
enum { ORC_DIFFUSE = 0, ORC_METAL = 1, ORC_DIELECTRIC = 2, ORC_EMISSION = 3 };
typedef struct {
    int32_t type;
    float   r, g, b;   /* Color (alpha is always 1.0, color.rs:21-23) */
    float   param;     /* Metal: fuzz; Dielectric: ir */
} orc_material;

/* common.rs:54-58 */
typedef struct { orc_vec3 center; float radius; orc_material material; } orc_sphere;
/* common.rs:101-107 */
typedef struct { orc_vec3 v0, v1, v2, normal; orc_material material; } orc_triangle;
/* camera.rs:8-15 */
typedef struct { orc_vec3 origin, lower_left_corner, horizontal, vertical; } orc_camera;

typedef struct orc_world orc_world;

/* RNG modes.  SERIAL is the reference's behaviour (one stream for the whole frame,
 * random.rs:8-10 + common.rs:321).  PER_SAMPLE is the counter-seeded stream the GPU
 * path uses: state(pixel, sample) = orc_sample_seed(seed, pixel, sample), then the
 * same xorshift32 update / f32 conversion as random.rs:15-30. */
enum { ORC_RNG_SERIAL = 0, ORC_RNG_PER_SAMPLE = 1 };

typedef struct {
    int32_t  samples_per_pixel;   /* common.rs:290 */
    int32_t  max_ray_bounces;     /* common.rs:291 */
    int32_t  rng_mode;            /* ORC_RNG_* */
    uint32_t seed;                /* 2547549 = random.rs:9 */
    int32_t  fixed_jitter;        /* !=0: sub-pixel offset (0.5,0.5), no jitter draws */
    int32_t  sample_begin;        /* PER_SAMPLE only: first sample index (progressive passes) */
    int32_t  threads;             /* PER_SAMPLE only: OpenMP threads over rows (<=1: serial) */
    int32_t  reserved;
} orc_options;

/* ---- L0 kernels exported for known-answer tests ---- */
uint32_t orc_xorshift32(uint32_t *state);                       /* random.rs:22-30 */
float    orc_random_f32(uint32_t *state);                       /* random.rs:15-17 */
float    orc_random_bilateral_f32(uint32_t *state);             /* random.rs:19-21 */
uint32_t orc_sample_seed(uint32_t seed, uint32_t pixel, uint32_t sample);
orc_vec3 orc_negate(orc_vec3 v);                                /* maths.rs:218-220 */
orc_vec3 orc_normalize(orc_vec3 v);                             /* maths.rs:111-118 */
orc_vec3 orc_cross(orc_vec3 a, orc_vec3 b);                     /* maths.rs:88-94 */
orc_vec3 orc_reflect(orc_vec3 v, orc_vec3 n);                   /* maths.rs:26-28 */
orc_vec3 orc_refract(orc_vec3 uv, orc_vec3 n, float ratio);     /* maths.rs:31-36 */
orc_vec3 orc_project(orc_vec3 v, orc_vec3 onto);                /* maths.rs:21-23 */
uint8_t  orc_f32_as_u8(float x);                                /* Rust `as u8` */

/* ---- camera.rs ---- */
orc_camera orc_camera_new_at(orc_vec3 origin, float aspect_ratio);                       /* :21-33 */
orc_camera orc_camera_new_with_vertical_fov(orc_vec3 origin, float vfov, float aspect);  /* :34-48 */
int        orc_camera_new_look_at(orc_vec3 origin, orc_vec3 look_at, orc_vec3 up,
                                  float vfov, float aspect, orc_camera *out);            /* :49-69 */
float      orc_camera_aspect_ratio(const orc_camera *c);                                 /* :70-72 */
orc_camera orc_move_camera_position(const orc_camera *c, float x, float y, float z);     /* lib.rs:60-63 */
void       orc_cast_ray(const orc_camera *c, float s, float t,
                        orc_vec3 *origin, orc_vec3 *direction);                          /* :84-89 */

/* ---- world ---- */
orc_world *orc_world_new(void);
void       orc_world_free(orc_world *w);
void       orc_world_add_sphere(orc_world *w, orc_vec3 c, float radius, orc_material m);
void       orc_world_add_triangle(orc_world *w, orc_vec3 v0, orc_vec3 v1, orc_vec3 v2,
                                  orc_material m);                                       /* common.rs:116-123 */
size_t     orc_world_sphere_count(const orc_world *w);
size_t     orc_world_triangle_count(const orc_world *w);
const orc_sphere   *orc_world_spheres(const orc_world *w);
const orc_triangle *orc_world_triangles(const orc_world *w);

/* parser.rs:336-382.  Returns NULL on error and sets *err (see orc_parse_error_name). */
enum { ORC_OK = 0, ORC_ERR_MISSING_CAMERA = 2, ORC_ERR_WRONG_SYNTAX = 3,
       ORC_ERR_DIDNT_START_WITH = 4, ORC_ERR_NOT_A_F32 = 6, ORC_ERR_UTF8 = 7 };
orc_world  *orc_parse_input(const char *source, orc_camera *camera_out, int *err);
const char *orc_parse_error_name(int err);

/* common.rs:237-258 — closest hit.  Returns 1 on hit.  prim_index: sphere i -> i,
 * triangle j -> n_spheres + j. */
int orc_world_hit(const orc_world *w, orc_vec3 origin, orc_vec3 direction,
                  float *t, orc_vec3 *position, orc_vec3 *normal, int64_t *prim_index);

/* common.rs:320-361.  pixels: width*height RGBA8, row-major, top row first.
 * accum_out (optional, may be NULL): width*height*4 floats, the un-resolved colour sums
 * indexed like pixels.  accum_in (optional): sums to continue from (progressive passes;
 * when NULL the accumulator starts at Color::new(0,0,0) = (0,0,0,1), common.rs:333).
 * resolve_spp: the divisor used in the resolve (normally == samples_per_pixel; the total
 * so far when accumulating progressively).  ray_count_out (optional): number of World::hit
 * calls (= ray segments).  Returns 0 on success. */
int orc_ray_trace(const orc_world *w, const orc_camera *camera,
                  uint8_t *pixels, size_t width, size_t height,
                  const orc_options *opt, int32_t resolve_spp,
                  const float *accum_in, float *accum_out, uint64_t *ray_count_out);

/* The same loop restricted to image rows [image_row_begin, image_row_end) (0 = top row), PER_SAMPLE
 * mode only: with one stream per (pixel, sample) a band is exactly that band of the full frame, which
 * is how full-width slices of the large BASELINE configs are checked in seconds.  Pixels and
 * accumulators outside the band are not touched; ray_count_out counts the band's segments. */
int orc_ray_trace_rows(const orc_world *w, const orc_camera *camera,
                       uint8_t *pixels, size_t width, size_t height,
                       size_t image_row_begin, size_t image_row_end,
                       const orc_options *opt, int32_t resolve_spp,
                       const float *accum_in, float *accum_out, uint64_t *ray_count_out);

/* image.rs:59-81 — ASCII PPM (P3).  Returns 0 on success. */
int orc_write_image(const uint8_t *pixels, size_t width, size_t height, const char *path);

#ifdef __cplusplus
}
#endif
#endif
