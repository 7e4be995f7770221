/*
 * rt_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see rt_oracle.h).
 *
 * Plain-C restatement of the reference's render loop.  Every function cites the
 * reference file:line (relative to /root/reference/raytracer/src/) it follows.
 * "image-level parity unpinned": see the header comment of rt_oracle.h.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math [-fopenmp] -fPIC -shared
 *   -ffp-contract=off : Rust never contracts a*b+c into an FMA; x86-64 SSE scalar
 *                       binary32 with correctly rounded / and sqrt is the reference
 *                       arithmetic.
 */
#define _GNU_SOURCE
#include "rt_oracle.h"
#include "unicode_alnum.h"

#include <locale.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ maths.rs */

typedef orc_vec3 v3;

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }

/* maths.rs:148 (add), :154 (sub), :160 (mul), :166 (div): component-wise in x,y,z */
static inline v3 v_add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v_sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
/* maths.rs:204-209: both `v * s` and `s * v` evaluate v.c * s */
static inline v3 v_muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
/* maths.rs:211-216 */
static inline v3 v_divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
/* maths.rs:218-220 */
static inline v3 v_neg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* maths.rs:82 / :125: x*x' + y*y' + z*z' (left-to-right) */
static inline float v_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* maths.rs:111-118: NVec3::new — three true divides, no zero guard */
static inline v3 v_normalize(v3 a)
{
    float length = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return V(a.x / length, a.y / length, a.z / length);
}
/* maths.rs:88-94 and :131-137 (same formula for Vec3 and NVec3) */
static inline v3 v_cross(v3 a, v3 b)
{
    return V(a.y * b.z - a.z * b.y, -(a.x * b.z - a.z * b.x), a.x * b.y - a.y * b.x);
}
/* maths.rs:46-49 */
static inline int v_near_zero(v3 a)
{
    const float s = 1e-8f;
    return (fabsf(a.x) < s) && (fabsf(a.y) < s) && (fabsf(a.z) < s);
}

orc_vec3 orc_negate(orc_vec3 v) { return v_neg(v); }
orc_vec3 orc_normalize(orc_vec3 v) { return v_normalize(v); }
orc_vec3 orc_cross(orc_vec3 a, orc_vec3 b) { return v_cross(a, b); }

/* maths.rs:26-28: v - 2.0 * v.dot(&n) * n   ==  v - ((2.0 * (v.n)) * n) */
orc_vec3 orc_reflect(orc_vec3 v, orc_vec3 n)
{
    return v_sub(v, v_muls(n, 2.0f * v_dot(v, n)));
}

/* maths.rs:31-36 */
orc_vec3 orc_refract(orc_vec3 uv, orc_vec3 n, float etai_over_etat)
{
    float cos_theta      = v_dot(v_neg(uv), n);
    v3    r_out_perp     = v_muls(v_add(uv, v_muls(n, cos_theta)), etai_over_etat);
    v3    r_out_parallel = v_muls(n, -sqrtf(fabsf(1.0f - v_dot(r_out_perp, r_out_perp))));
    return v_add(r_out_perp, r_out_parallel);
}

/* maths.rs:21-23 (not on the hot path; restated only for its known-answer test) */
orc_vec3 orc_project(orc_vec3 v, orc_vec3 onto)
{
    return v_muls(onto, v_dot(onto, v) / v_dot(onto, onto));
}

/* ----------------------------------------------------------------- random.rs */

/* random.rs:22-30 */
uint32_t orc_xorshift32(uint32_t *state)
{
    uint32_t x = *state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    *state = x;
    return x;
}
/* random.rs:15-17: `u32::MAX as f32` rounds to 2^32 */
float orc_random_f32(uint32_t *state)
{
    return (float)orc_xorshift32(state) / 4294967296.0f;
}
/* random.rs:19-21 */
float orc_random_bilateral_f32(uint32_t *state)
{
    return orc_random_f32(state) * 2.0f - 1.0f;
}

/* Per-(pixel, sample) stream seed of the parallel path.  No reference counterpart: the
 * reference has one serial stream (common.rs:321).  Contract (DESIGN.md "RNG"):
 * lowbias32 finaliser twice; never returns 0 (xorshift fixed point, random.rs:11). */
static inline uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
uint32_t orc_sample_seed(uint32_t seed, uint32_t pixel, uint32_t sample)
{
    uint32_t h = mix32(pixel ^ seed);
    h = mix32(h + sample * 0x9E3779B9U + 0x85EBCA6BU);
    return h ? h : 0x9E3779B9U;
}

/* common.rs:32-38: NVec3::new(b, b, b) — normalised cube sample, draws in x,y,z order */
static inline v3 random_unit_sphere(uint32_t *rng)
{
    float x = orc_random_bilateral_f32(rng);
    float y = orc_random_bilateral_f32(rng);
    float z = orc_random_bilateral_f32(rng);
    return v_normalize(V(x, y, z));
}

/* ------------------------------------------------------------------ color.rs */

typedef struct { float r, g, b, a; } color;
static inline color C(float r, float g, float b) { color c = { r, g, b, 1.0f }; return c; } /* :21-23 */
/* color.rs:30 */
static inline color c_add_with_alpha(color a, color b)
{
    color c = { a.r + b.r, a.g + b.g, a.b + b.b, a.a + b.a };
    return c;
}
/* color.rs:36 */
static inline color c_mul_with_alpha(color a, color b)
{
    color c = { a.r * b.r, a.g * b.g, a.b * b.b, a.a * b.a };
    return c;
}

/* Rust `f32 as u8`: truncate toward zero, saturate, NaN -> 0 */
uint8_t orc_f32_as_u8(float x)
{
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 255.0f) return 255;
    return (uint8_t)x;
}

/* ----------------------------------------------------------------- camera.rs */

/* camera.rs:21-33 */
orc_camera orc_camera_new_at(orc_vec3 origin, float aspect_ratio)
{
    float viewport_height = 2.0f;
    float viewport_width  = aspect_ratio * viewport_height;
    float focal_length    = 1.0f;
    orc_camera c;
    c.origin     = origin;
    c.horizontal = V(viewport_width, 0.0f, 0.0f);
    c.vertical   = V(0.0f, viewport_height, 0.0f);
    c.lower_left_corner =
        v_sub(origin, V(viewport_width / 2.0f, viewport_height / 2.0f, focal_length));
    return c;
}

/* camera.rs:34-48 */
orc_camera orc_camera_new_with_vertical_fov(orc_vec3 origin, float vfov, float aspect_ratio)
{
    float h               = tanf(vfov / 2.0f);
    float viewport_height = 2.0f * h;
    float viewport_width  = aspect_ratio * viewport_height;
    float focal_length    = 1.0f;
    orc_camera c;
    c.origin     = origin;
    c.horizontal = V(viewport_width, 0.0f, 0.0f);
    c.vertical   = V(0.0f, viewport_height, 0.0f);
    c.lower_left_corner =
        v_sub(origin, V(viewport_width / 2.0f, viewport_height / 2.0f, focal_length));
    return c;
}

/* camera.rs:49-69.  `up` is an NVec3 in the reference, i.e. already normalised by
 * construction (maths.rs:111-118); a raw triple is normalised here the same way.
 * Returns 0 on success, 1/2 for the two asserts (:50, :62). */
int orc_camera_new_look_at(orc_vec3 origin, orc_vec3 look_at, orc_vec3 up_raw,
                           float vfov, float aspect_ratio, orc_camera *out)
{
    if (v_near_zero(v_sub(origin, look_at))) return 1;
    float viewport_height = 2.0f * tanf(vfov / 2.0f);
    float viewport_width  = viewport_height * aspect_ratio;

    v3 up = v_normalize(up_raw);
    v3 w  = v_normalize(v_sub(origin, look_at));
    v3 u  = v_cross(up, w);   /* NVec3::cross = new_unchecked: NOT normalised */
    v3 v  = v_cross(w, u);
    if (!(fabsf(v.y) > 1e-8f)) return 2;

    v3 horizontal = v_muls(u, viewport_width);
    v3 vertical   = v_muls(v, viewport_height);
    out->origin     = origin;
    out->horizontal = horizontal;
    out->vertical   = vertical;
    out->lower_left_corner =
        v_sub(v_sub(v_sub(origin, v_divs(horizontal, 2.0f)), v_divs(vertical, 2.0f)), w);
    return 0;
}

/* camera.rs:70-72 */
float orc_camera_aspect_ratio(const orc_camera *c) { return c->horizontal.x / c->vertical.y; }

/* lib.rs:60-63 */
orc_camera orc_move_camera_position(const orc_camera *c, float x, float y, float z)
{
    return orc_camera_new_at(v_add(c->origin, V(x, y, z)), orc_camera_aspect_ratio(c));
}

/* camera.rs:84-89: ((llc + s*horizontal) + t*vertical) - origin, normalised */
void orc_cast_ray(const orc_camera *c, float s, float t, orc_vec3 *origin, orc_vec3 *direction)
{
    v3 p = v_sub(v_add(v_add(c->lower_left_corner, v_muls(c->horizontal, s)),
                       v_muls(c->vertical, t)),
                 c->origin);
    *origin    = c->origin;
    *direction = v_normalize(p);
}

/* ----------------------------------------------------------------- common.rs */

struct orc_world {
    orc_sphere   *spheres;   size_t n_spheres, cap_spheres;
    orc_triangle *triangles; size_t n_triangles, cap_triangles;   /* the single Mesh, lib.rs:41 */
};

orc_world *orc_world_new(void) { return (orc_world *)calloc(1, sizeof(orc_world)); }
void orc_world_free(orc_world *w)
{
    if (!w) return;
    free(w->spheres);
    free(w->triangles);
    free(w);
}
void orc_world_add_sphere(orc_world *w, orc_vec3 c, float radius, orc_material m)
{
    if (w->n_spheres == w->cap_spheres) {
        w->cap_spheres = w->cap_spheres ? 2 * w->cap_spheres : 16;
        w->spheres = (orc_sphere *)realloc(w->spheres, w->cap_spheres * sizeof(orc_sphere));
    }
    orc_sphere s = { c, radius, m };
    w->spheres[w->n_spheres++] = s;
}
/* common.rs:116-123: stored normal = normalize((v1-v0) x (v2-v0)) */
void orc_world_add_triangle(orc_world *w, orc_vec3 v0, orc_vec3 v1, orc_vec3 v2, orc_material m)
{
    if (w->n_triangles == w->cap_triangles) {
        w->cap_triangles = w->cap_triangles ? 2 * w->cap_triangles : 16;
        w->triangles =
            (orc_triangle *)realloc(w->triangles, w->cap_triangles * sizeof(orc_triangle));
    }
    v3 a = v_sub(v1, v0);
    v3 b = v_sub(v2, v0);
    orc_triangle t = { v0, v1, v2, v_normalize(v_cross(a, b)), m };
    w->triangles[w->n_triangles++] = t;
}
size_t orc_world_sphere_count(const orc_world *w) { return w->n_spheres; }
size_t orc_world_triangle_count(const orc_world *w) { return w->n_triangles; }
const orc_sphere   *orc_world_spheres(const orc_world *w) { return w->spheres; }
const orc_triangle *orc_world_triangles(const orc_world *w) { return w->triangles; }

typedef struct { v3 origin, direction; } ray;
typedef struct { v3 position, normal; float t; const orc_material *material; } hit_record;

/* common.rs:20 */
static inline v3 ray_at(const ray *r, float t) { return v_add(r->origin, v_muls(r->direction, t)); }

/* common.rs:60-98 */
static int sphere_hit(const orc_sphere *s, const ray *r, float t_min, float t_max, hit_record *out)
{
    v3    oc     = v_sub(r->origin, s->center);
    float a      = 1.0f;                               /* NVec3::length_squared, maths.rs:127 */
    float half_b = v_dot(oc, r->direction);
    float c      = v_dot(oc, oc) - s->radius * s->radius;   /* powi(2) */
    float discriminant = half_b * half_b - a * c;

    if (discriminant < 0.0f) return 0;

    float discriminant_sqrt = sqrtf(discriminant);
    float root1 = (-half_b - discriminant_sqrt) / a;
    float root2 = (-half_b + discriminant_sqrt) / a;

    /* :88-92: filter (t_min < x < t_max) then min_by */
    int   have = 0;
    float t    = 0.0f;
    if (t_min < root1 && root1 < t_max) { t = root1; have = 1; }
    if (t_min < root2 && root2 < t_max) {
        if (!have || root2 < t) { t = root2; have = 1; }
    }
    if (!have) return 0;

    out->t        = t;
    out->position = ray_at(r, t);
    out->normal   = v_normalize(v_divs(v_sub(out->position, s->center), s->radius));
    out->material = &s->material;
    return 1;
}

/* common.rs:124-166 */
static int triangle_intersect(const orc_triangle *tr, const ray *r, float t_min, float t_max,
                              hit_record *out)
{
    v3 v0 = tr->v0, v1 = tr->v1, v2 = tr->v2;
    v3 a = v_sub(v1, v0);
    v3 b = v_sub(v2, v0);
    v3 n = v_cross(a, b);

    float cos_angle_and_length = v_dot(n, r->direction);
    if (-1e-8f < cos_angle_and_length && cos_angle_and_length < 1e-8f) return 0;

    float d = v_dot(n, v0);
    float t = (v_dot(n, r->origin) + d) / cos_angle_and_length;      /* sic, :140-141 */
    if (t < t_min || t > t_max) return 0;

    v3 p = ray_at(r, t);

    v3 e0 = v_sub(v1, v0), vp0 = v_sub(p, v0);
    if (v_dot(n, v_cross(e0, vp0)) < 0.0f) return 0;
    v3 e1 = v_sub(v2, v1), vp1 = v_sub(p, v1);
    if (v_dot(n, v_cross(e1, vp1)) < 0.0f) return 0;
    v3 e2 = v_sub(v0, v2), vp2 = v_sub(p, v2);
    if (v_dot(n, v_cross(e2, vp2)) < 0.0f) return 0;

    out->position = p;
    out->normal   = tr->normal;
    out->t        = t;
    out->material = &tr->material;
    return 1;
}

/* common.rs:178-223 */
static int mesh_hit(const orc_world *w, const ray *r, float t_min, float t_max, hit_record *out,
                    int64_t *tri_index)
{
    int   found = 0;
    float closest_intersection = INFINITY;
    for (size_t i = 0; i < w->n_triangles; ++i) {
        hit_record h;
        if (triangle_intersect(&w->triangles[i], r, t_min, t_max, &h)) {
            if (h.t < closest_intersection) {
                closest_intersection = h.t;
                *out       = h;
                *tri_index = (int64_t)i;
                found      = 1;
            }
        }
    }
    return found;
}

/* common.rs:237-258 */
static int world_hit(const orc_world *w, const ray *r, hit_record *out, int64_t *prim)
{
    float closest = INFINITY;
    int   found   = 0;
    for (size_t i = 0; i < w->n_spheres; ++i) {
        hit_record h;
        if (sphere_hit(&w->spheres[i], r, 0.001f, closest, &h)) {
            closest = h.t;
            *out    = h;
            if (prim) *prim = (int64_t)i;
            found = 1;
        }
    }
    /* exactly one Mesh (lib.rs:41, main.rs:58-79), possibly empty */
    {
        hit_record h;
        int64_t    ti = -1;
        if (mesh_hit(w, r, 0.001f, closest, &h, &ti)) {
            closest = h.t;
            *out    = h;
            if (prim) *prim = (int64_t)w->n_spheres + ti;
            found = 1;
        }
    }
    return found;
}

int orc_world_hit(const orc_world *w, orc_vec3 origin, orc_vec3 direction, float *t,
                  orc_vec3 *position, orc_vec3 *normal, int64_t *prim_index)
{
    ray        r = { origin, direction };
    hit_record h;
    int64_t    prim = -1;
    if (!world_hit(w, &r, &h, &prim)) return 0;
    if (t) *t = h.t;
    if (position) *position = h.position;
    if (normal) *normal = h.normal;
    if (prim_index) *prim_index = prim;
    return 1;
}

/* -------------------------------------------------------------- materials.rs */

typedef struct { color col; int has_next; ray next; } scatter_data;

/* materials.rs:26-28 (name is inverted in the reference; semantics kept) */
static inline int hit_front_face(v3 direction, v3 normal) { return v_dot(direction, normal) >= 0.0f; }

/* materials.rs:42-52 */
static scatter_data diffuse_scatter(const orc_material *m, const hit_record *hit, uint32_t *rng)
{
    scatter_data s;
    v3 scatter = v_add(hit->normal, random_unit_sphere(rng));
    s.col      = C(m->r, m->g, m->b);
    s.has_next = 1;
    s.next.origin = hit->position;
    s.next.direction = v_near_zero(scatter) ? hit->normal : v_normalize(scatter);
    return s;
}

/* materials.rs:54-63 */
static scatter_data metal_scatter(const orc_material *m, const ray *r, const hit_record *hit,
                                  uint32_t *rng)
{
    scatter_data s;
    v3 reflected = orc_reflect(r->direction, hit->normal);
    v3 direction = v_add(reflected, v_muls(random_unit_sphere(rng), m->param));
    s.col = C(m->r, m->g, m->b);
    if (hit_front_face(direction, hit->normal)) {
        s.has_next = 1;
        s.next.origin    = hit->position;
        s.next.direction = v_normalize(direction);
    } else {
        s.has_next = 0;
    }
    return s;
}

/* materials.rs:65-97 */
static scatter_data dielectric_scatter(const orc_material *m, const ray *r, const hit_record *hit)
{
    scatter_data s;
    float ir = m->param;
    v3    normal;
    float refraction_ratio;
    if (hit_front_face(r->direction, hit->normal)) {
        normal = v_neg(hit->normal);
        refraction_ratio = 1.0f / ir;
    } else {
        normal = hit->normal;
        refraction_ratio = ir;
    }
    v3 refracted = orc_refract(r->direction, normal, refraction_ratio);
    s.col      = C(1.0f, 1.0f, 1.0f);
    s.has_next = 1;
    s.next.origin    = hit->position;
    s.next.direction = v_normalize(refracted);
    return s;
}

/* materials.rs:31-39, :100-102 */
static scatter_data material_scatter(const orc_material *m, const ray *r, const hit_record *hit,
                                     uint32_t *rng)
{
    switch (m->type) {
    case ORC_DIFFUSE:    return diffuse_scatter(m, hit, rng);
    case ORC_METAL:      return metal_scatter(m, r, hit, rng);
    case ORC_DIELECTRIC: return dielectric_scatter(m, r, hit);
    default: {
        scatter_data s;
        s.col = C(m->r, m->g, m->b);
        s.has_next = 0;
        return s;
    }
    }
}

/* common.rs:263-285 */
static color ray_color(const ray *ray_in, const orc_world *w, uint32_t *rng, int depth,
                       uint64_t *segments)
{
    ray   r = *ray_in;
    color final_color = C(1.0f, 1.0f, 1.0f);

    for (int i = 0; i < depth; ++i) {
        hit_record hit;
        ++*segments;
        if (world_hit(w, &r, &hit, NULL)) {
            scatter_data s = material_scatter(hit.material, &r, &hit, rng);
            if (s.has_next) {
                final_color = c_mul_with_alpha(final_color, s.col);
                r = s.next;
            } else {
                return c_mul_with_alpha(final_color, s.col);
            }
        } else {
            /* :278-280: lerp(a,b,t) = a*(1.0-t) + b*t on Vec3, common.rs:26-29 */
            float t  = 0.5f * (v_normalize(r.direction).y + 1.0f);
            v3    sk = v_add(v_muls(V(1.0f, 1.0f, 1.0f), 1.0f - t), v_muls(V(0.5f, 0.7f, 1.0f), t));
            return c_mul_with_alpha(final_color, C(sk.x, sk.y, sk.z));
        }
    }
    return C(0.0f, 0.0f, 0.0f);   /* :284: Vec3::new_zero().into() — alpha 1 */
}

/* One pixel of common.rs:332-357. */
static void trace_pixel(const orc_world *w, const orc_camera *cam, size_t width, size_t height,
                        size_t row, size_t column, const orc_options *opt, int32_t resolve_spp,
                        uint32_t *serial_rng, const float *accum_in, float *accum_out,
                        uint8_t *pixels, uint64_t *segments)
{
    size_t out_index = (height - row - 1) * width + column;      /* :351 */
    color  col = { 0.0f, 0.0f, 0.0f, 1.0f };                     /* :333 Color::new(0,0,0) */
    if (accum_in) {
        col.r = accum_in[4 * out_index + 0];
        col.g = accum_in[4 * out_index + 1];
        col.b = accum_in[4 * out_index + 2];
        col.a = accum_in[4 * out_index + 3];
    }
    for (int32_t s = 0; s < opt->samples_per_pixel; ++s) {
        uint32_t  local;
        uint32_t *rng = serial_rng;
        if (opt->rng_mode == ORC_RNG_PER_SAMPLE) {
            local = orc_sample_seed(opt->seed, (uint32_t)(row * width + column),
                                    (uint32_t)(opt->sample_begin + s));
            rng = &local;
        }
        float ju = opt->fixed_jitter ? 0.5f : orc_random_f32(rng);   /* :335, u first */
        float u  = ((float)column + ju) / (float)(width - 1);
        float jv = opt->fixed_jitter ? 0.5f : orc_random_f32(rng);   /* :336 */
        float v  = ((float)row + jv) / (float)(height - 1);
        ray r;
        orc_cast_ray(cam, u, v, &r.origin, &r.direction);
        col = c_add_with_alpha(col, ray_color(&r, w, rng, opt->max_ray_bounces, segments));
    }
    if (accum_out) {
        accum_out[4 * out_index + 0] = col.r;
        accum_out[4 * out_index + 1] = col.g;
        accum_out[4 * out_index + 2] = col.b;
        accum_out[4 * out_index + 3] = col.a;
    }
    /* :344-356 */
    float k = 1.0f / (float)resolve_spp;
    pixels[4 * out_index + 0] = orc_f32_as_u8(sqrtf(col.r * k) * 255.999f);
    pixels[4 * out_index + 1] = orc_f32_as_u8(sqrtf(col.g * k) * 255.999f);
    pixels[4 * out_index + 2] = orc_f32_as_u8(sqrtf(col.b * k) * 255.999f);
    pixels[4 * out_index + 3] = orc_f32_as_u8(col.a * k * 255.999f);
}

/* common.rs:320-361 restricted to the image rows [image_row_begin, image_row_end) of the frame
 * (image row 0 = top; the reference's loop row is height-1-image_row, :351).  Only meaningful in
 * PER_SAMPLE mode, where every (pixel, sample) has its own stream and a band is therefore exactly
 * the band of the full frame.  Pixels / accumulators outside the band are left untouched. */
int orc_ray_trace_rows(const orc_world *w, const orc_camera *camera, uint8_t *pixels, size_t width,
                       size_t height, size_t image_row_begin, size_t image_row_end,
                       const orc_options *opt, int32_t resolve_spp,
                       const float *accum_in, float *accum_out, uint64_t *ray_count_out)
{
    uint64_t total = 0;
    if (opt->rng_mode != ORC_RNG_PER_SAMPLE) return 1;
    if (image_row_end > height) image_row_end = height;
    if (image_row_begin > image_row_end) image_row_begin = image_row_end;
    int threads = opt->threads > 1 ? opt->threads : 1;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total)
#endif
    for (long ir = (long)image_row_begin; ir < (long)image_row_end; ++ir) {
        uint64_t segs = 0;
        size_t   row  = height - 1 - (size_t)ir;
        for (size_t column = 0; column < width; ++column)
            trace_pixel(w, camera, width, height, row, column, opt, resolve_spp, NULL,
                        accum_in, accum_out, pixels, &segs);
        total += segs;
    }
    if (ray_count_out) *ray_count_out = total;
    return 0;
}

/* common.rs:320-361 */
int orc_ray_trace(const orc_world *w, const orc_camera *camera, uint8_t *pixels, size_t width,
                  size_t height, const orc_options *opt, int32_t resolve_spp,
                  const float *accum_in, float *accum_out, uint64_t *ray_count_out)
{
    uint64_t total = 0;
    if (opt->rng_mode == ORC_RNG_SERIAL) {
        uint32_t rng = opt->seed;                                 /* :321 */
        for (size_t row = 0; row < height; ++row)                 /* :327 */
            for (size_t column = 0; column < width; ++column)     /* :332 */
                trace_pixel(w, camera, width, height, row, column, opt, resolve_spp, &rng,
                            accum_in, accum_out, pixels, &total);
    } else if (opt->rng_mode == ORC_RNG_PER_SAMPLE) {
        int threads = opt->threads > 1 ? opt->threads : 1;
        (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total)
#endif
        for (long row = 0; row < (long)height; ++row) {
            uint64_t segs = 0;
            for (size_t column = 0; column < width; ++column)
                trace_pixel(w, camera, width, height, (size_t)row, column, opt, resolve_spp, NULL,
                            accum_in, accum_out, pixels, &segs);
            total += segs;
        }
    } else {
        return 1;
    }
    if (ray_count_out) *ray_count_out = total;
    return 0;
}

/* ------------------------------------------------------------------ image.rs */

/* image.rs:59-81 */
int orc_write_image(const uint8_t *pixels, size_t width, size_t height, const char *path)
{
    FILE *f = path ? fopen(path, "w") : stdout;
    if (!f) return 1;
    fprintf(f, "P3\n%zu %zu\n%d\n", width, height, 255);
    for (size_t row = 0; row < height; ++row)
        for (size_t column = 0; column < width; ++column) {
            const uint8_t *p = &pixels[4 * (row * width + column)];
            fprintf(f, "%u %u %u\n", p[0], p[1], p[2]);
        }
    if (path) fclose(f);
    return 0;
}

/* ----------------------------------------------------------------- parser.rs */

typedef struct { const char *p; size_t n; } str;     /* a &str slice */

/* Decode one UTF-8 scalar; returns its byte length (input is validated up front). */
static size_t utf8_decode(const char *s, size_t n, uint32_t *cp)
{
    const unsigned char *u = (const unsigned char *)s;
    if (u[0] < 0x80 || n < 2) { *cp = u[0]; return 1; }
    if ((u[0] & 0xE0) == 0xC0) { *cp = ((u[0] & 0x1Fu) << 6) | (u[1] & 0x3Fu); return 2; }
    if ((u[0] & 0xF0) == 0xE0 && n >= 3) {
        *cp = ((u[0] & 0x0Fu) << 12) | ((u[1] & 0x3Fu) << 6) | (u[2] & 0x3Fu);
        return 3;
    }
    if (n >= 4) {
        *cp = ((u[0] & 0x07u) << 18) | ((u[1] & 0x3Fu) << 12) | ((u[2] & 0x3Fu) << 6) | (u[3] & 0x3Fu);
        return 4;
    }
    *cp = u[0];
    return 1;
}

/* Rust char::is_whitespace = Unicode White_Space */
static int is_unicode_whitespace(uint32_t c)
{
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 ||
           (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F ||
           c == 0x205F || c == 0x3000;
}

/* Strict UTF-8 validation (CStr::to_str, lib.rs:40). */
static int utf8_valid(const char *s, size_t n)
{
    const unsigned char *u = (const unsigned char *)s;
    size_t i = 0;
    while (i < n) {
        unsigned char c = u[i];
        size_t len;
        uint32_t cp;
        if (c < 0x80) { ++i; continue; }
        else if (c >= 0xC2 && c <= 0xDF) { len = 2; cp = c & 0x1Fu; }
        else if (c >= 0xE0 && c <= 0xEF) { len = 3; cp = c & 0x0Fu; }
        else if (c >= 0xF0 && c <= 0xF4) { len = 4; cp = c & 0x07u; }
        else return 0;
        if (i + len > n) return 0;
        for (size_t k = 1; k < len; ++k) {
            if ((u[i + k] & 0xC0) != 0x80) return 0;
            cp = (cp << 6) | (u[i + k] & 0x3Fu);
        }
        if (len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return 0;
        if (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)) return 0;
        i += len;
    }
    return 1;
}

/* parser.rs:54-57 */
static str skip_whitespace(str s)
{
    while (s.n) {
        uint32_t cp;
        size_t   len = utf8_decode(s.p, s.n, &cp);
        if (!is_unicode_whitespace(cp)) break;
        s.p += len; s.n -= len;
    }
    return s;
}

/* parser.rs:59-62: the longest prefix of chars for which char::is_alphanumeric() holds (Unicode
 * Alphabetic or general category N*, table in unicode_alnum.h) or that are '_'. */
static str get_identifier(str s, str *name)
{
    size_t i = 0;
    while (i < s.n) {
        unsigned char c = (unsigned char)s.p[i];
        if (c < 0x80) {
            if ((c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_') { ++i; continue; }
            break;
        }
        uint32_t cp;
        size_t   len = utf8_decode(s.p + i, s.n - i, &cp);
        if (!rt_is_unicode_alnum(cp)) break;
        i += len;
    }
    name->p = s.p; name->n = i;
    s.p += i; s.n -= i;
    return s;
}

/* parser.rs:82-89: 0 on success */
static int starts_with(str *s, const char *target)
{
    size_t size = strlen(target);
    if (s->n >= size && memcmp(s->p, target, size) == 0) { s->p += size; s->n -= size; return 0; }
    return ORC_ERR_DIDNT_START_WITH;
}

static locale_t c_locale(void)
{
    static locale_t loc = (locale_t)0;
    if (!loc) loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    return loc;
}

/* parser.rs:107-133 */
static int parse_float(str *s, float *out)
{
    const char *data = s->p;
    int    found_dot = 0;
    size_t index = 0, digits = 0;

    if (s->n < 3) return ORC_ERR_NOT_A_F32;                 /* :112-114 */
    if (data[0] == '-') index = 1;
    while (index < s->n) {
        char c = data[index];
        if (c >= '0' && c <= '9') { ++index; ++digits; }
        else if (c == '.') {
            if (found_dot) return ORC_ERR_NOT_A_F32;
            found_dot = 1; ++index;
        } else break;
    }
    /* str::parse::<f32> accepts "5", "5.", ".5", "-.5"; rejects "", "-", ".", "-." */
    if (digits == 0) return ORC_ERR_NOT_A_F32;
    char  stackbuf[64];
    char *buf = index < sizeof stackbuf ? stackbuf : (char *)malloc(index + 1);
    memcpy(buf, data, index);
    buf[index] = 0;
    *out = strtof_l(buf, NULL, c_locale());                  /* correctly rounded, like Rust */
    if (buf != stackbuf) free(buf);
    s->p += index; s->n -= index;
    return 0;
}

/* parser.rs:135-142 */
static int parse_vec3(str *s, v3 *out)
{
    int e;
    if ((e = parse_float(s, &out->x))) return e;
    *s = skip_whitespace(*s);
    if ((e = parse_float(s, &out->y))) return e;
    *s = skip_whitespace(*s);
    if ((e = parse_float(s, &out->z))) return e;
    return 0;
}

#define TRY(expr) do { int e_ = (expr); if (e_) return e_; } while (0)
#define WS(s) ((s) = skip_whitespace(s))

/* parser.rs:145-167.  Returns -1 when the statement keyword is absent (Rust: None). */
static int parse_camera(str *src, orc_camera *cam)
{
    str s = *src;
    if (starts_with(&s, "camera")) return -1;
    v3 o; float a;
    WS(s); TRY(starts_with(&s, "origin")); WS(s); TRY(parse_vec3(&s, &o)); WS(s);
    TRY(starts_with(&s, "aspect")); WS(s); TRY(parse_float(&s, &a)); WS(s);
    TRY(starts_with(&s, ";"));
    *cam = orc_camera_new_at(o, a);
    *src = s;
    return 0;
}

/* parser.rs:175-234 */
static int parse_material(str *src, str *name, orc_material *m)
{
    str s = *src;
    if (starts_with(&s, "material")) return -1;
    WS(s);
    s = get_identifier(s, name);
    WS(s); TRY(starts_with(&s, ":")); WS(s);

    str t = s;
    if (starts_with(&t, "Diffuse") == 0) {
        v3 c;
        WS(t); TRY(starts_with(&t, "color")); WS(t); TRY(parse_vec3(&t, &c)); WS(t);
        TRY(starts_with(&t, ";"));
        m->type = ORC_DIFFUSE; m->r = c.x; m->g = c.y; m->b = c.z; m->param = 0.0f;
        *src = t;
        return 0;
    }
    t = s;
    if (starts_with(&t, "Metal") == 0) {
        v3 c; float f;
        WS(t); TRY(starts_with(&t, "color")); WS(t); TRY(parse_vec3(&t, &c)); WS(t);
        TRY(starts_with(&t, "fuzz")); WS(t); TRY(parse_float(&t, &f)); WS(t);
        TRY(starts_with(&t, ";"));
        m->type = ORC_METAL; m->r = c.x; m->g = c.y; m->b = c.z; m->param = f;
        *src = t;
        return 0;
    }
    t = s;
    if (starts_with(&t, "Dielectric") == 0) {
        float i;
        WS(t); TRY(starts_with(&t, "ir")); WS(t); TRY(parse_float(&t, &i)); WS(t);
        TRY(starts_with(&t, ";"));
        m->type = ORC_DIELECTRIC; m->r = 1.0f; m->g = 1.0f; m->b = 1.0f; m->param = i;
        *src = t;
        return 0;
    }
    return ORC_ERR_WRONG_SYNTAX;
}

typedef struct { str name; orc_material m; } named_material;
typedef struct { named_material *v; size_t n, cap; } material_map;

/* HashMap::insert (parser.rs:355): a later definition of the same name replaces the earlier */
static void map_insert(material_map *map, str name, orc_material m)
{
    for (size_t i = 0; i < map->n; ++i)
        if (map->v[i].name.n == name.n && memcmp(map->v[i].name.p, name.p, name.n) == 0) {
            map->v[i].m = m;
            return;
        }
    if (map->n == map->cap) {
        map->cap = map->cap ? 2 * map->cap : 16;
        map->v = (named_material *)realloc(map->v, map->cap * sizeof(named_material));
    }
    map->v[map->n].name = name;
    map->v[map->n].m = m;
    ++map->n;
}
static const orc_material *map_get(const material_map *map, str name)
{
    for (size_t i = 0; i < map->n; ++i)
        if (map->v[i].name.n == name.n && memcmp(map->v[i].name.p, name.p, name.n) == 0)
            return &map->v[i].m;
    return NULL;
}

/* parser.rs:237-269 */
static int parse_sphere(str *src, const material_map *map, orc_world *w)
{
    str s = *src;
    if (starts_with(&s, "sphere")) return -1;
    v3 c; float r; str m;
    WS(s); TRY(starts_with(&s, "center")); WS(s); TRY(parse_vec3(&s, &c)); WS(s);
    TRY(starts_with(&s, "radius")); WS(s); TRY(parse_float(&s, &r)); WS(s);
    TRY(starts_with(&s, "material")); WS(s); s = get_identifier(s, &m); WS(s);
    TRY(starts_with(&s, ";"));
    const orc_material *mat = map_get(map, m);
    if (!mat) return ORC_ERR_WRONG_SYNTAX;
    orc_world_add_sphere(w, c, r, *mat);
    *src = s;
    return 0;
}

/* parser.rs:272-310 */
static int parse_triangle(str *src, const material_map *map, orc_world *w)
{
    str s = *src;
    if (starts_with(&s, "triangle")) return -1;
    v3 v0, v1, v2; str m;
    WS(s); TRY(starts_with(&s, "v0")); WS(s); TRY(parse_vec3(&s, &v0)); WS(s);
    TRY(starts_with(&s, "v1")); WS(s); TRY(parse_vec3(&s, &v1)); WS(s);
    TRY(starts_with(&s, "v2")); WS(s); TRY(parse_vec3(&s, &v2)); WS(s);
    TRY(starts_with(&s, "material")); WS(s); s = get_identifier(s, &m); WS(s);
    TRY(starts_with(&s, ";"));
    const orc_material *mat = map_get(map, m);
    if (!mat) return ORC_ERR_WRONG_SYNTAX;
    orc_world_add_triangle(w, v0, v1, v2, *mat);
    *src = s;
    return 0;
}

/* parser.rs:313-323: a comment must end in '\n'; nothing is skipped after it */
static int skip_comment(str *s)
{
    for (;;) {
        str t = *s;
        if (starts_with(&t, "//")) return 0;
        const char *nl = (const char *)memchr(t.p, '\n', t.n);
        if (!nl) return ORC_ERR_WRONG_SYNTAX;
        size_t adv = (size_t)(nl - t.p) + 1;
        s->p = t.p + adv; s->n = t.n - adv;
    }
}

static int parse_input_impl(str s, orc_camera *cam, orc_world *w, material_map *map)
{
    int e;
    TRY(skip_comment(&s));                                     /* :342 */
    e = parse_camera(&s, cam);                                 /* :343-350 */
    if (e == -1) return ORC_ERR_MISSING_CAMERA;
    if (e) return e;
    WS(s);
    TRY(skip_comment(&s));                                     /* :353 */
    for (;;) {                                                 /* :354-359 */
        str name; orc_material m;
        e = parse_material(&s, &name, &m);
        if (e == -1) break;
        if (e) return e;
        map_insert(map, name, m);
        WS(s); TRY(skip_comment(&s));
    }
    for (;;) {                                                 /* :362-367 */
        e = parse_sphere(&s, map, w);
        if (e == -1) break;
        if (e) return e;
        WS(s); TRY(skip_comment(&s));
    }
    for (;;) {                                                 /* :370-375 */
        e = parse_triangle(&s, map, w);
        if (e == -1) break;
        if (e) return e;
        WS(s); TRY(skip_comment(&s));
    }
    return s.n ? ORC_ERR_WRONG_SYNTAX : 0;                     /* :377-381 */
}

/* parser.rs:336-382 (+ lib.rs:39-40 for the NUL-terminated UTF-8 input) */
orc_world *orc_parse_input(const char *source, orc_camera *camera_out, int *err)
{
    str s = { source, strlen(source) };
    int e;
    if (!utf8_valid(s.p, s.n)) { if (err) *err = ORC_ERR_UTF8; return NULL; }
    orc_world   *w = orc_world_new();
    material_map map = { 0, 0, 0 };
    orc_camera   cam;
    memset(&cam, 0, sizeof cam);
    e = parse_input_impl(s, &cam, w, &map);
    free(map.v);
    if (err) *err = e;
    if (e) { orc_world_free(w); return NULL; }
    if (camera_out) *camera_out = cam;
    return w;
}

const char *orc_parse_error_name(int err)
{
    switch (err) {
    case ORC_OK: return "Ok";
    case ORC_ERR_MISSING_CAMERA: return "MissingCamera";
    case ORC_ERR_WRONG_SYNTAX: return "WrongSyntax";
    case ORC_ERR_DIDNT_START_WITH: return "DidntStartWith";
    case ORC_ERR_NOT_A_F32: return "NotAF32";
    case ORC_ERR_UTF8: return "Utf8Error";
    default: return "Error";
    }
}
