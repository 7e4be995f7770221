// rt_scene.cpp — host-side scene construction for the render path: camera constructors,
// triangle precomputation, SoA packing (kernel K1 of SURVEY.md §2a is a single H2D copy of
// the blob built here), the world-text parser that load_world() needs, and the PPM writer.
//
// Compiled with -ffp-contract=off: every value computed here (camera vectors, r*r, triangle
// normals, plane constants) must carry exactly the bits the reference's Rust code computes.
#include "rt_host.hpp"
#include "rt_unicode_alnum.h"

#include <clocale>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <locale.h>
#include <string_view>
#include <unordered_map>

namespace rt {

namespace {

struct F3 { float x, y, z; };
inline F3 f3(RtVec3 v) { return {v.x, v.y, v.z}; }
inline RtVec3 rv(F3 v) { return {v.x, v.y, v.z}; }
inline F3 sub(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 add(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline F3 scale(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline F3 divide(F3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// maths.rs:88-94
inline F3 cross3(F3 a, F3 b) { return {a.y * b.z - a.z * b.y, -(a.x * b.z - a.z * b.x), a.x * b.y - a.y * b.x}; }
// maths.rs:111-118
inline F3 unit(F3 a)
{
    float len = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return {a.x / len, a.y / len, a.z / len};
}
inline bool tiny(F3 a) { return std::fabs(a.x) < 1e-8f && std::fabs(a.y) < 1e-8f && std::fabs(a.z) < 1e-8f; }

}   // namespace

// ------------------------------------------------------------------ camera.rs

Camera Camera::new_at(RtVec3 origin, float aspect_ratio)
{
    const float viewport_height = 2.0f;
    const float viewport_width  = aspect_ratio * viewport_height;
    const float focal_length    = 1.0f;
    Camera c;
    c.d.origin            = origin;
    c.d.horizontal        = {viewport_width, 0.0f, 0.0f};
    c.d.vertical          = {0.0f, viewport_height, 0.0f};
    c.d.lower_left_corner = rv(sub(f3(origin), F3{viewport_width / 2.0f, viewport_height / 2.0f, focal_length}));
    return c;
}

Camera Camera::new_with_vertical_fov(RtVec3 origin, float vfov, float aspect_ratio)
{
    const float h               = std::tan(vfov / 2.0f);
    const float viewport_height = 2.0f * h;
    const float viewport_width  = aspect_ratio * viewport_height;
    const float focal_length    = 1.0f;
    Camera c;
    c.d.origin            = origin;
    c.d.horizontal        = {viewport_width, 0.0f, 0.0f};
    c.d.vertical          = {0.0f, viewport_height, 0.0f};
    c.d.lower_left_corner = rv(sub(f3(origin), F3{viewport_width / 2.0f, viewport_height / 2.0f, focal_length}));
    return c;
}

bool Camera::new_look_at(RtVec3 origin, RtVec3 look_at, RtVec3 up_in, float vfov, float aspect_ratio,
                         Camera* out, std::string* error)
{
    if (tiny(sub(f3(origin), f3(look_at)))) {                       // camera.rs:50
        if (error) *error = "Origin and look_at must differ!";
        return false;
    }
    const float viewport_height = 2.0f * std::tan(vfov / 2.0f);
    const float viewport_width  = viewport_height * aspect_ratio;

    const F3 up = unit(f3(up_in));                                  // `up: NVec3` is unit by construction
    const F3 w  = unit(sub(f3(origin), f3(look_at)));
    const F3 u  = cross3(up, w);                                    // NVec3::cross: not re-normalised
    const F3 v  = cross3(w, u);
    if (!(std::fabs(v.y) > 1e-8f)) {                                // camera.rs:62
        if (error) *error = "Origin and look_at can't have the same z-coordinate.";
        return false;
    }
    const F3 horizontal = scale(u, viewport_width);
    const F3 vertical   = scale(v, viewport_height);
    out->d.origin       = origin;
    out->d.horizontal   = rv(horizontal);
    out->d.vertical     = rv(vertical);
    out->d.lower_left_corner =
        rv(sub(sub(sub(f3(origin), divide(horizontal, 2.0f)), divide(vertical, 2.0f)), w));
    return true;
}

Camera Camera::moved(float x, float y, float z) const
{
    return Camera::new_at(rv(add(f3(d.origin), F3{x, y, z})), aspect_ratio());
}

// ----------------------------------------------------------------- common.rs

Triangle Triangle::make(RtVec3 v0, RtVec3 v1, RtVec3 v2, const Material& m)
{
    const F3 a = sub(f3(v1), f3(v0));
    const F3 b = sub(f3(v2), f3(v0));
    Triangle t;
    t.v0 = v0; t.v1 = v1; t.v2 = v2;
    t.normal   = rv(unit(cross3(a, b)));
    t.material = m;
    return t;
}

std::unique_ptr<World> World::make(std::vector<Sphere> spheres, std::vector<Triangle> triangles)
{
    auto w = std::make_unique<World>();
    w->spheres   = std::move(spheres);
    w->triangles = std::move(triangles);
    return w;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

RtSceneView World::Packed::view(const unsigned char* base) const
{
    RtSceneView v{};
    v.sph       = reinterpret_cast<const RtFloat4*>(base + off_sph);
    v.tri_plane = reinterpret_cast<const RtFloat4*>(base + off_tri_plane);
    v.sph_filter = reinterpret_cast<const RtFloat4*>(base + off_sph_filter);
    v.sph_r2     = reinterpret_cast<const float*>(base + off_sph_r2);
    v.cull_bound = reinterpret_cast<const RtFloat4*>(base + off_cull_bound);
    v.cull_sph   = reinterpret_cast<const RtFloat4*>(base + off_cull_sph);
    v.cull_r2    = reinterpret_cast<const float*>(base + off_cull_r2);
    v.cull_orig  = reinterpret_cast<const uint32_t*>(base + off_cull_orig);
    v.n_groups   = n_groups;
    v.tri_cull  = reinterpret_cast<const RtFloat4*>(base + off_tri_cull);
    v.tri_v     = reinterpret_cast<const RtFloat4*>(base + off_tri_v);
    v.info      = reinterpret_cast<const RtPrimInfo*>(base + off_info);
    v.n_sph     = n_sph;
    v.n_sph_pad = n_sph_pad;
    v.n_tri     = n_tri;
    v.n_tri_pad = n_tri_pad;
    return v;
}

namespace {
RtPrimInfo prim_info(const Material& m, float radius)
{
    RtPrimInfo i{};
    i.r = m.r; i.g = m.g; i.b = m.b;
    i.param     = m.param;
    i.type      = (uint32_t)m.type;
    i.inv_param = 1.0f / m.param;        // materials.rs:69 `1.0 / ir`, the same IEEE divide
    i.radius    = radius;
    return i;
}
}   // namespace

// Conservative edge-stage reject data of one triangle (rt_trace.cuh, triangle_group).
// The reference's three inside tests (common.rs:147-163) are  n.(E_k x (p - v_k)) >= 0  with
// n = (v1-v0)x(v2-v0); divided by |n|^2 these are the barycentric coordinates of p's projection
// onto the triangle's plane:  lambda2 = f0/|n|^2,  lambda0 = f1/|n|^2,  lambda1 = f2/|n|^2, each
// an affine function  G.p + g  of p.  The kernel evaluates lambda2 and lambda0 approximately and
// rejects p when one coordinate is below -0.5 (lambda1 = 1 - lambda0 - lambda2), i.e. when p is
// outside the triangle by at least half of the corresponding height — provided the operands are
// small enough for every rounding error (of this approximation AND of the reference's own
// float evaluation, <= 12u |p - v_k| / h_k) to stay below 0.02: |o|_1 + t + |v0|_1 < 2^12 h_min,
// which the kernel checks as  |o|_1 + t < K.  Degenerate, needle-thin (K <= 0) or non-finite
// triangles get K = -inf and are never rejected.  Evaluated in double, rounded once.
namespace {
void triangle_cull_record(const Triangle& t, RtFloat4 out[3])
{
    const double v0[3] = {t.v0.x, t.v0.y, t.v0.z}, v1[3] = {t.v1.x, t.v1.y, t.v1.z}, v2[3] = {t.v2.x, t.v2.y, t.v2.z};
    auto subd   = [](const double a[3], const double b[3], double r[3]) { for (int i = 0; i < 3; ++i) r[i] = a[i] - b[i]; };
    auto crossd = [](const double a[3], const double b[3], double r[3]) {
        r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0]; };
    auto dotd   = [](const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    double e0[3], eb[3], e1[3], n[3], g2[3], g0[3];
    subd(v1, v0, e0); subd(v2, v0, eb); subd(v2, v1, e1);
    crossd(e0, eb, n);
    const double nn = dotd(n, n);
    const float  ninf = -INFINITY;
    out[0] = {0.f, 0.f, 0.f, 0.f}; out[1] = {0.f, 0.f, 0.f, 0.f}; out[2] = {ninf, 0.f, 0.f, 0.f};
    if (!(nn > 0.0) || !std::isfinite(nn)) return;
    crossd(n, e0, g2);                       // f0 = n.(e0 x (p - v0)) = (n x e0).(p - v0)
    crossd(n, e1, g0);                       // f1 = n.(e1 x (p - v1)) = (n x e1).(p - v1)
    for (int i = 0; i < 3; ++i) { g2[i] /= nn; g0[i] /= nn; }
    const double c2 = -dotd(g2, v0), c0 = -dotd(g0, v1);
    // heights: h_k = |n| / |edge_k|
    double e2[3]; subd(v0, v2, e2);
    const double ln = std::sqrt(nn);
    const double lmax = std::sqrt(std::max(dotd(e0, e0), std::max(dotd(e1, e1), dotd(e2, e2))));
    const double hmin = ln / lmax;
    const double a0 = std::fabs(v0[0]) + std::fabs(v0[1]) + std::fabs(v0[2]);
    const double a1 = std::fabs(v1[0]) + std::fabs(v1[1]) + std::fabs(v1[2]);
    const double a2 = std::fabs(v2[0]) + std::fabs(v2[1]) + std::fabs(v2[2]);
    const double K = 4096.0 * hmin - std::max(a0, std::max(a1, a2));
    out[0] = {(float)g2[0], (float)g2[1], (float)g2[2], (float)c2};
    out[1] = {(float)g0[0], (float)g0[1], (float)g0[2], (float)c0};
    const bool finite = std::isfinite(out[0].x) && std::isfinite(out[0].y) && std::isfinite(out[0].z) && std::isfinite(out[0].w) &&
                        std::isfinite(out[1].x) && std::isfinite(out[1].y) && std::isfinite(out[1].z) && std::isfinite(out[1].w);
    if (finite && K > 0.0 && std::isfinite(K)) out[2].x = std::nextafterf((float)K, ninf);
}
}   // namespace

// ---- block C: spatial order and group bounds of the CULL kernels (rt_trace.cuh, cull_spheres) ----
// Spheres much larger than the typical one (a ground sphere) would blow up the bound of any
// group they sit in, so they come first, in groups of their own that always pass; the rest
// are sorted along a Morton curve through their centres and cut into groups of 8.
namespace {
struct CullOrder { std::vector<std::vector<uint32_t>> groups; size_t always = 0; };   // groups[g] = list indices

uint32_t morton3(uint32_t x, uint32_t y, uint32_t z)
{
    auto spread = [](uint32_t v) { v &= 0x3ffu; v = (v | (v << 16)) & 0x30000ffu; v = (v | (v << 8)) & 0x300f00fu;
                                   v = (v | (v << 4)) & 0x30c30c3u; v = (v | (v << 2)) & 0x9249249u; return v; };
    return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}

CullOrder cull_order(const std::vector<Sphere>& sph)
{
    CullOrder o;
    const size_t S = sph.size();
    std::vector<float> radii(S);
    for (size_t i = 0; i < S; ++i) radii[i] = std::fabs(sph[i].radius);
    std::vector<float> sorted = radii;
    std::nth_element(sorted.begin(), sorted.begin() + S / 2, sorted.end());
    const float big = 8.0f * sorted[S / 2];
    std::vector<uint32_t> large, small;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = 0; i < S; ++i) {
        const Sphere& s = sph[i];
        const bool finite = std::isfinite(s.center.x) && std::isfinite(s.center.y) && std::isfinite(s.center.z) &&
                            std::isfinite(s.radius);
        if (!finite || !(radii[i] <= big)) { large.push_back((uint32_t)i); continue; }
        small.push_back((uint32_t)i);
        const double c[3] = {s.center.x, s.center.y, s.center.z};
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], c[a]); hi[a] = std::max(hi[a], c[a]); }
    }
    std::vector<std::pair<uint32_t, uint32_t>> keyed(small.size());
    for (size_t k = 0; k < small.size(); ++k) {
        const Sphere& s = sph[small[k]];
        const double c[3] = {s.center.x, s.center.y, s.center.z};
        uint32_t q[3];
        for (int a = 0; a < 3; ++a) {
            const double ext = hi[a] - lo[a];
            q[a] = ext > 0 ? (uint32_t)std::min(1023.0, (c[a] - lo[a]) / ext * 1023.0) : 0u;
        }
        keyed[k] = {morton3(q[0], q[1], q[2]), small[k]};
    }
    std::sort(keyed.begin(), keyed.end());
    for (size_t k = 0; k < large.size(); k += 8)
        o.groups.emplace_back(large.begin() + k, large.begin() + std::min(k + 8, large.size()));
    o.always = o.groups.size();
    for (size_t k = 0; k < keyed.size(); k += 8) {
        std::vector<uint32_t> g;
        for (size_t j = k; j < std::min(k + 8, keyed.size()); ++j) g.push_back(keyed[j].second);
        o.groups.push_back(std::move(g));
    }
    return o;
}

// Bound of a group (rt_trace.cuh cull_spheres gives the derivation): a member can pass its filter
// only if the ray's line passes within  R(o) = A + B|o|  of the group centre cB, where
//   A = Rgeo + B*Cg,  Rgeo = max_k(|c_k - cB| + r_k),  Cg = max_k(|c_k| + r_k),  B = RT_CULL_B.
// The kernel evaluates  (o.d - cB.d)^2 - (o.o - 2 cB.o + cB.cB) + R(o)^2  + margins as
//   fma(hb,hb,-t) + fma(gB, |o|, -kray),   t = -2 cB.o + wB,
//   wB = cB.cB - A^2 - m (cB.cB + A^2)  (rounded down),   gB = 2AB (1 + m)  (rounded up).
void build_cull_block(const CullOrder& order, const RtFloat4* sph_filter, const float* sph_r2, RtFloat4* bound,
                      RtFloat4* sph9, float* r2, uint32_t* orig)
{
    const float  nan  = std::nanf("");
    const float  ninf = -INFINITY;
    const double B = (double)RT_CULL_B, m = (double)RT_CULL_M;
    for (size_t g = 0; g < order.groups.size(); ++g) {
        const std::vector<uint32_t>& idx = order.groups[g];
        if (idx.empty()) {                                     // padding group: never passes, never hit
            bound[g] = RtFloat4{nan, nan, nan, nan};
            for (size_t k = 0; k < 8; ++k) { sph9[9 * g + k] = RtFloat4{nan, nan, nan, nan}; r2[8 * g + k] = nan; orig[8 * g + k] = 0xffffffffu; }
            sph9[9 * g + 8] = RtFloat4{0.f, 0.f, 0.f, 0.f};
            continue;
        }
        double cb[3] = {0, 0, 0};
        for (uint32_t i : idx) { cb[0] += sph_filter[i].x; cb[1] += sph_filter[i].y; cb[2] += sph_filter[i].z; }
        for (int a = 0; a < 3; ++a) cb[a] /= (double)idx.size();
        const float cbf[3] = {(float)cb[0], (float)cb[1], (float)cb[2]};          // the centre the kernel uses
        double rgeo = 0.0, cg = 0.0;
        for (uint32_t i : idx) {
            const double c[3] = {sph_filter[i].x, sph_filter[i].y, sph_filter[i].z};
            const double r = std::sqrt((double)sph_r2[i]);
            const double dx = c[0] - cbf[0], dy = c[1] - cbf[1], dz = c[2] - cbf[2];
            rgeo = std::max(rgeo, std::sqrt(dx * dx + dy * dy + dz * dz) + r);
            cg   = std::max(cg, std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) + r);
        }
        rgeo = rgeo * (1.0 + 1e-6) + 1e-30;
        const double A   = rgeo + B * cg;
        const double ccb = (double)cbf[0] * cbf[0] + (double)cbf[1] * cbf[1] + (double)cbf[2] * cbf[2];
        const double w   = ccb - A * A - m * (ccb + A * A) - 1e-30;
        const double gb  = 2.0 * A * B * (1.0 + m);
        float wf = (float)w, gf = (float)gb;
        if ((double)wf > w) wf = std::nextafterf(wf, ninf);
        if ((double)gf < gb) gf = std::nextafterf(gf, INFINITY);
        const bool usable = g >= order.always && std::isfinite(wf) && std::isfinite(gf) && std::isfinite(cbf[0]) &&
                            std::isfinite(cbf[1]) && std::isfinite(cbf[2]);
        bound[g] = usable ? RtFloat4{cbf[0], cbf[1], cbf[2], wf} : RtFloat4{0.f, 0.f, 0.f, ninf};   // -inf: always passes
        for (size_t k = 0; k < 8; ++k) {
            const bool real = k < idx.size();
            sph9[9 * g + k] = real ? sph_filter[idx[k]] : RtFloat4{nan, nan, nan, nan};
            r2[8 * g + k]   = real ? sph_r2[idx[k]] : nan;
            orig[8 * g + k] = real ? idx[k] : 0xffffffffu;
        }
        sph9[9 * g + 8] = RtFloat4{usable ? gf : 0.f, 0.f, 0.f, 0.f};
    }
}
}   // namespace

// Scene pack: AoS {Sphere, Triangle} -> the SoA blob of rt_types.h.
const World::Packed& World::packed() const
{
    if (packed_) return *packed_;
    auto p = std::make_unique<Packed>();
    const size_t S = spheres.size(), T = triangles.size(), P = S + T;
    const size_t Sp = align_up(S, S >= RT_FILTER_FROM ? RT_FILTER_GROUP : RT_SPHERE_GROUP);
    p->n_sph     = (uint32_t)S;
    p->n_sph_pad = (uint32_t)Sp;
    const size_t Tp = align_up(T, RT_TRI_GROUP);
    p->n_tri     = (uint32_t)T;
    p->n_tri_pad = (uint32_t)Tp;
    size_t off = 0;
    p->off_sph       = off; off += Sp * sizeof(RtFloat4);
    p->off_tri_plane = off; off += Tp * sizeof(RtFloat4);
    p->off_sph_filter = off; off += Sp * sizeof(RtFloat4);
    const size_t off_plane_b = off; off += Tp * sizeof(RtFloat4);
    p->off_sph_r2    = off; off += align_up(Sp * sizeof(float), 16);
    // block C exists for the sphere counts the FILTER kernels serve; order and bounds first
    CullOrder order;
    if (S >= RT_FILTER_FROM) order = cull_order(spheres);
    while (order.groups.size() % 32u) order.groups.emplace_back();     // whole rounds of 32 groups (empty = NaN groups)
    const size_t Gc = order.groups.size();
    p->n_groups = (uint32_t)Gc;
    p->off_cull_bound = off; off += Gc * sizeof(RtFloat4);
    p->off_cull_sph   = off; off += 9 * Gc * sizeof(RtFloat4);
    const size_t off_plane_c = off; off += Tp * sizeof(RtFloat4);
    p->off_cull_r2    = off; off += align_up(8 * Gc * sizeof(float), 16);
    p->off_cull_orig  = off; off += align_up(8 * Gc * sizeof(uint32_t), 16);
    p->off_tri_cull  = off; off += 5 * (Tp / 2) * sizeof(RtFloat4);
    p->off_tri_v     = off; off += 3 * T * sizeof(RtFloat4);
    off = align_up(off, 32);
    p->off_info      = off; off += P * sizeof(RtPrimInfo);
    p->blob.assign(off ? off : 32, 0);
    unsigned char* base = p->blob.data();
    auto* sph   = reinterpret_cast<RtFloat4*>(base + p->off_sph);
    auto* plane = reinterpret_cast<RtFloat4*>(base + p->off_tri_plane);
    auto* triv  = reinterpret_cast<RtFloat4*>(base + p->off_tri_v);
    auto* cull  = reinterpret_cast<RtFloat4*>(base + p->off_tri_cull);
    auto* info  = reinterpret_cast<RtPrimInfo*>(base + p->off_info);
    for (size_t i = 0; i < S; ++i) {
        const Sphere& s = spheres[i];
        sph[i]  = {s.center.x, s.center.y, s.center.z, s.radius * s.radius};   // radius.powi(2), common.rs:77
        info[i] = prim_info(s.material, s.radius);
    }
    const float nan = std::nanf("");
    for (size_t i = S; i < Sp; ++i) sph[i] = {nan, nan, nan, nan};            // padding: never hit
    std::vector<RtFloat4> sph_aos(sph, sph + Sp);                             // {c, r*r} per sphere, for the lists below
    std::vector<RtFloat4> plane_aos(Tp, RtFloat4{nan, nan, nan, nan});       // {n, n.v0} per triangle; NaN padding
    std::vector<RtFloat4> cull_aos(3 * Tp, RtFloat4{0.f, 0.f, 0.f, 0.f});
    for (size_t j = T; j < Tp; ++j) cull_aos[3 * j + 2].x = -INFINITY;        // padding: never culled, never hit (NaN plane)
    // block B: the conservative filter list of the exact kernel (rt_trace.cuh, sphere_filter_group)
    // The filter records {c, w} are computed per sphere (also the input of block C) and stored for the kernels
    // in PAIRS, two float4 per pair of consecutive spheres: {x0, x1, y0, y1} {z0, z1, -w0, -w1} — the operand
    // layout of the two-wide FMAs (FFMA2) the filter runs on.  Sp is even (a multiple of 8).
    std::vector<RtFloat4> sphf_aos(Sp);
    RtFloat4* sphf = sphf_aos.data();
    auto* pairs = reinterpret_cast<RtFloat4*>(base + p->off_sph_filter);
    auto* r2   = reinterpret_cast<float*>(base + p->off_sph_r2);
    for (size_t i = 0; i < Sp; ++i) {
        const RtFloat4 si = sph_aos[i];
        sphf[i] = si;
        r2[i]   = si.w;
        // w = c.c - r^2 - 2^-17 (c.c + r^2), in double, rounded DOWN (a smaller w only lets more spheres through)
        const double cc = (double)si.x * si.x + (double)si.y * si.y + (double)si.z * si.z;
        const double rr = (double)si.w;
        const double w  = cc - rr - (cc + rr) * (1.0 / 131072.0) - 1e-30;
        float wf = (float)w;
        if ((double)wf > w) wf = std::nextafterf(wf, -INFINITY);
        sphf[i].w = wf;                                            // NaN padding stays NaN
    }
    for (size_t i = 0; i + 1 < Sp; i += 2) {
        pairs[i]     = {sphf[i].x, sphf[i + 1].x, sphf[i].y, sphf[i + 1].y};
        pairs[i + 1] = {sphf[i].z, sphf[i + 1].z, -sphf[i].w, -sphf[i + 1].w};
        // block A, the direct kernels' list, in the same pair layout with r*r in the last slots
        sph[i]       = {sph_aos[i].x, sph_aos[i + 1].x, sph_aos[i].y, sph_aos[i + 1].y};
        sph[i + 1]   = {sph_aos[i].z, sph_aos[i + 1].z, sph_aos[i].w, sph_aos[i + 1].w};
    }
    for (size_t j = 0; j < T; ++j) {
        const Triangle& t = triangles[j];
        // common.rs:128-133,140: n = (v1-v0) x (v2-v0) and d = n.v0 depend on the triangle
        // only, so they are evaluated once here with the reference's exact operation order.
        const F3 n = cross3(sub(f3(t.v1), f3(t.v0)), sub(f3(t.v2), f3(t.v0)));
        plane_aos[j]    = {n.x, n.y, n.z, dot3(n, f3(t.v0))};
        triv[3 * j + 0] = {t.v0.x, t.v0.y, t.v0.z, t.normal.x};
        triv[3 * j + 1] = {t.v1.x, t.v1.y, t.v1.z, t.normal.y};
        triv[3 * j + 2] = {t.v2.x, t.v2.y, t.v2.z, t.normal.z};
        info[S + j]     = prim_info(t.material, 1.0f);
        triangle_cull_record(t, &cull_aos[3 * j]);
    }
    // the hot / warm triangle lists in PAIRS of consecutive triangles — the operand layout of the two-wide FP32
    // instructions of rt_trace.cuh triangle_group_n (Tp is even; the last pair of an odd list ends in the padding)
    for (size_t j = 0; j + 1 < Tp; j += 2) {
        const RtFloat4 a = plane_aos[j], b = plane_aos[j + 1];
        plane[j]     = {a.x, b.x, a.y, b.y};
        plane[j + 1] = {a.z, b.z, a.w, b.w};
        const RtFloat4 *ca = &cull_aos[3 * j], *cb = &cull_aos[3 * (j + 1)];
        RtFloat4* q = &cull[5 * (j / 2)];
        q[0] = {ca[0].x, cb[0].x, ca[0].y, cb[0].y};
        q[1] = {ca[0].z, cb[0].z, ca[0].w, cb[0].w};
        q[2] = {ca[1].x, cb[1].x, ca[1].y, cb[1].y};
        q[3] = {ca[1].z, cb[1].z, ca[1].w, cb[1].w};
        q[4] = {ca[2].x, cb[2].x, 0.f, 0.f};
    }
    std::memcpy(base + off_plane_b, plane, Tp * sizeof(RtFloat4));
    std::memcpy(base + off_plane_c, plane, Tp * sizeof(RtFloat4));
    if (Gc) build_cull_block(order, sphf,
                             reinterpret_cast<const float*>(base + p->off_sph_r2),
                             reinterpret_cast<RtFloat4*>(base + p->off_cull_bound),
                             reinterpret_cast<RtFloat4*>(base + p->off_cull_sph),
                             reinterpret_cast<float*>(base + p->off_cull_r2),
                             reinterpret_cast<uint32_t*>(base + p->off_cull_orig));
    packed_ = std::move(p);
    return *packed_;
}

// ----------------------------------------------------------------- world -> text
// The inverse of parse_input: the world in the reference's grammar (parser.rs:326-335).
// parse_float accepts only `-?digits[.digits]` (no exponent), so every f32 is printed as its
// exact finite decimal expansion, which any correctly rounding parser reads back bit for bit.
namespace {
void put_float(std::string& out, float v)
{
    if (!std::isfinite(v)) v = 0.0f;                       // not expressible in the grammar
    char buf[512];
    int n = std::snprintf(buf, sizeof buf, "%.160f", (double)v);   // exact: a float has < 150 fractional digits
    while (n > 0 && buf[n - 1] == '0') --n;
    if (n > 0 && buf[n - 1] == '.') { buf[n++] = '0'; }
    if (n < 3) { buf[n++] = '0'; }                         // parse_float needs >= 3 bytes of input (parser.rs:112-114)
    out.append(buf, (size_t)n);
}
void put_vec(std::string& out, RtVec3 v) { put_float(out, v.x); out += ' '; put_float(out, v.y); out += ' '; put_float(out, v.z); }
void put_material(std::string& out, size_t index, const Material& m)
{
    out += "material M" + std::to_string(index) + " : ";
    switch (m.type) {
    case RT_MAT_DIFFUSE:    out += "Diffuse color "; put_vec(out, RtVec3{m.r, m.g, m.b}); break;
    case RT_MAT_METAL:      out += "Metal color "; put_vec(out, RtVec3{m.r, m.g, m.b}); out += " fuzz "; put_float(out, m.param); break;
    case RT_MAT_DIELECTRIC: out += "Dielectric ir "; put_float(out, m.param); break;
    default:                out += "Emission color "; put_vec(out, RtVec3{m.r, m.g, m.b}); break;
    }
    out += ";\n";
}
}   // namespace

std::string world_to_text(const World& w, const Camera& camera)
{
    std::string out;
    out.reserve(128 * (w.spheres.size() + w.triangles.size()) + 256);
    // the grammar's camera is Camera::new_at(origin, aspect) (parser.rs:145-167)
    out += "camera origin "; put_vec(out, camera.d.origin); out += " aspect "; put_float(out, camera.aspect_ratio()); out += ";\n";
    for (size_t i = 0; i < w.spheres.size(); ++i) put_material(out, i, w.spheres[i].material);
    for (size_t j = 0; j < w.triangles.size(); ++j) put_material(out, w.spheres.size() + j, w.triangles[j].material);
    for (size_t i = 0; i < w.spheres.size(); ++i) {
        out += "sphere center "; put_vec(out, w.spheres[i].center); out += " radius "; put_float(out, w.spheres[i].radius);
        out += " material M" + std::to_string(i) + ";\n";
    }
    for (size_t j = 0; j < w.triangles.size(); ++j) {
        const Triangle& t = w.triangles[j];
        out += "triangle v0 "; put_vec(out, t.v0); out += " v1 "; put_vec(out, t.v1); out += " v2 "; put_vec(out, t.v2);
        out += " material M" + std::to_string(w.spheres.size() + j) + ";\n";
    }
    return out;
}

// ----------------------------------------------------------------- parser.rs

const char* parse_error_name(ParseError e)
{
    switch (e) {
    case ParseError::Ok: return "Ok";
    case ParseError::CouldntOpenFile: return "Couldn't open file";
    case ParseError::MissingCamera: return "Missing camera";
    case ParseError::WrongSyntax: return "Wrong syntax";
    case ParseError::DidntStartWith: return "Error. (DidntStartWith)";
    case ParseError::NotAI32: return "Error. (NotAI32)";
    case ParseError::NotAF32: return "Error. (NotAF32)";
    case ParseError::InvalidUtf8: return "Invalid UTF-8";
    }
    return "Error.";
}

namespace {

using sv = std::string_view;

struct Fail { ParseError e; };   // thrown inside the parser only; never crosses parse_input

// One Unicode scalar from valid UTF-8.
inline size_t decode(sv s, uint32_t& cp)
{
    auto u = [&](size_t i) { return (uint32_t)(unsigned char)s[i]; };
    if (u(0) < 0x80) { cp = u(0); return 1; }
    if ((u(0) & 0xE0) == 0xC0 && s.size() >= 2) { cp = ((u(0) & 0x1F) << 6) | (u(1) & 0x3F); return 2; }
    if ((u(0) & 0xF0) == 0xE0 && s.size() >= 3) { cp = ((u(0) & 0x0F) << 12) | ((u(1) & 0x3F) << 6) | (u(2) & 0x3F); return 3; }
    if (s.size() >= 4) { cp = ((u(0) & 0x07) << 18) | ((u(1) & 0x3F) << 12) | ((u(2) & 0x3F) << 6) | (u(3) & 0x3F); return 4; }
    cp = u(0);
    return 1;
}

// char::is_whitespace (Unicode White_Space)
inline bool is_space(uint32_t c)
{
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 ||
           (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}

bool valid_utf8(sv s)
{
    size_t i = 0, n = s.size();
    while (i < n) {
        unsigned char c = (unsigned char)s[i];
        size_t len; uint32_t cp;
        if (c < 0x80) { ++i; continue; }
        if (c >= 0xC2 && c <= 0xDF) { len = 2; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { len = 3; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { len = 4; cp = c & 0x07; }
        else return false;
        if (i + len > n) return false;
        for (size_t k = 1; k < len; ++k) {
            unsigned char d = (unsigned char)s[i + k];
            if ((d & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (d & 0x3F);
        }
        if (len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
        if (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
        i += len;
    }
    return true;
}

struct Cursor {
    sv s;

    void skip_whitespace()                                  // parser.rs:54-57
    {
        while (!s.empty()) {
            uint32_t cp; size_t len = decode(s, cp);
            if (!is_space(cp)) break;
            s.remove_prefix(len);
        }
    }
    // parser.rs:59-62: the longest prefix of chars with char::is_alphanumeric() (Unicode Alphabetic or N*) or '_'
    sv identifier()
    {
        size_t i = 0;
        while (i < s.size()) {
            const unsigned char c = (unsigned char)s[i];
            if (c < 0x80) {
                if ((c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_') { ++i; continue; }
                break;
            }
            uint32_t cp; const size_t len = decode(s.substr(i), cp);
            if (!rt_is_unicode_alnum(cp)) break;
            i += len;
        }
        sv name = s.substr(0, i);
        s.remove_prefix(i);
        return name;
    }
    bool accept(sv target)                                  // parser.rs:82-89 as a predicate
    {
        if (s.size() >= target.size() && s.compare(0, target.size(), target) == 0) {
            s.remove_prefix(target.size());
            return true;
        }
        return false;
    }
    void expect(sv target) { if (!accept(target)) throw Fail{ParseError::DidntStartWith}; }

    float number()                                          // parser.rs:107-133
    {
        if (s.size() < 3) throw Fail{ParseError::NotAF32};  // :112-114 (remaining *input* < 3 bytes)
        size_t index = (s[0] == '-') ? 1 : 0, digits = 0;
        bool   dot   = false;
        for (; index < s.size(); ++index) {
            char c = s[index];
            if (c >= '0' && c <= '9') ++digits;
            else if (c == '.') { if (dot) throw Fail{ParseError::NotAF32}; dot = true; }
            else break;
        }
        if (!digits) throw Fail{ParseError::NotAF32};       // "", "-", ".", "-." fail str::parse::<f32>
        std::string tok(s.substr(0, index));
        static locale_t c_loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
        float v = strtof_l(tok.c_str(), nullptr, c_loc);    // correctly rounded, as Rust's parse
        s.remove_prefix(index);
        return v;
    }
    RtVec3 vec3()                                           // parser.rs:135-142
    {
        RtVec3 v;
        v.x = number(); skip_whitespace();
        v.y = number(); skip_whitespace();
        v.z = number();
        return v;
    }
    void skip_comment()                                     // parser.rs:313-323
    {
        while (s.size() >= 2 && s[0] == '/' && s[1] == '/') {
            size_t nl = s.find('\n', 2);
            if (nl == sv::npos) throw Fail{ParseError::WrongSyntax};
            s.remove_prefix(nl + 1);
        }
    }
};

struct SvHash { size_t operator()(sv v) const { return std::hash<sv>{}(v); } };
using MaterialMap = std::unordered_map<sv, Material, SvHash>;

bool statement_camera(Cursor& c, Camera& cam)               // parser.rs:145-167
{
    if (!c.accept("camera")) return false;
    c.skip_whitespace(); c.expect("origin"); c.skip_whitespace();
    RtVec3 o = c.vec3(); c.skip_whitespace();
    c.expect("aspect"); c.skip_whitespace();
    float a = c.number(); c.skip_whitespace();
    c.expect(";");
    cam = Camera::new_at(o, a);
    return true;
}

bool statement_material(Cursor& c, MaterialMap& map, bool allow_emission)   // parser.rs:175-234
{
    if (!c.accept("material")) return false;
    c.skip_whitespace();
    sv name = c.identifier();
    c.skip_whitespace(); c.expect(":"); c.skip_whitespace();
    Material m;
    if (c.accept("Diffuse")) {
        c.skip_whitespace(); c.expect("color"); c.skip_whitespace();
        RtVec3 col = c.vec3(); c.skip_whitespace();
        c.expect(";");
        m = Material::Diffuse(col.x, col.y, col.z);
    } else if (c.accept("Metal")) {
        c.skip_whitespace(); c.expect("color"); c.skip_whitespace();
        RtVec3 col = c.vec3(); c.skip_whitespace();
        c.expect("fuzz"); c.skip_whitespace();
        float f = c.number(); c.skip_whitespace();
        c.expect(";");
        m = Material::Metal(col.x, col.y, col.z, f);
    } else if (c.accept("Dielectric")) {
        c.skip_whitespace(); c.expect("ir"); c.skip_whitespace();
        float ir = c.number(); c.skip_whitespace();
        c.expect(";");
        m = Material::Dielectric(ir);
    } else if (allow_emission && c.accept("Emission")) {
        // Opt-in extension (rt_load_world_ext): MaterialType::Emission exists
        // (materials.rs:11) but the reference grammar cannot express it (parser.rs:171-174) —
        // load_world keeps rejecting it, as the reference does; same shape as Diffuse.
        c.skip_whitespace(); c.expect("color"); c.skip_whitespace();
        RtVec3 col = c.vec3(); c.skip_whitespace();
        c.expect(";");
        m = Material::Emission(col.x, col.y, col.z);
    } else {
        throw Fail{ParseError::WrongSyntax};
    }
    map[name] = m;                                          // HashMap::insert: last definition wins
    return true;
}

const Material& lookup(const MaterialMap& map, sv name)
{
    auto it = map.find(name);
    if (it == map.end()) throw Fail{ParseError::WrongSyntax};   // parser.rs:259, :299
    return it->second;
}

bool statement_sphere(Cursor& c, const MaterialMap& map, std::vector<Sphere>& out)   // parser.rs:237-269
{
    if (!c.accept("sphere")) return false;
    c.skip_whitespace(); c.expect("center"); c.skip_whitespace();
    RtVec3 center = c.vec3(); c.skip_whitespace();
    c.expect("radius"); c.skip_whitespace();
    float r = c.number(); c.skip_whitespace();
    c.expect("material"); c.skip_whitespace();
    sv name = c.identifier(); c.skip_whitespace();
    c.expect(";");
    out.push_back(Sphere{center, r, lookup(map, name)});
    return true;
}

bool statement_triangle(Cursor& c, const MaterialMap& map, std::vector<Triangle>& out)   // parser.rs:272-310
{
    if (!c.accept("triangle")) return false;
    c.skip_whitespace(); c.expect("v0"); c.skip_whitespace();
    RtVec3 v0 = c.vec3(); c.skip_whitespace();
    c.expect("v1"); c.skip_whitespace();
    RtVec3 v1 = c.vec3(); c.skip_whitespace();
    c.expect("v2"); c.skip_whitespace();
    RtVec3 v2 = c.vec3(); c.skip_whitespace();
    c.expect("material"); c.skip_whitespace();
    sv name = c.identifier(); c.skip_whitespace();
    c.expect(";");
    out.push_back(Triangle::make(v0, v1, v2, lookup(map, name)));
    return true;
}

}   // namespace

ParseResult parse_input(const char* source, size_t length, bool allow_emission)
{
    ParseResult res;
    sv text(source, length);
    if (!valid_utf8(text)) { res.error = ParseError::InvalidUtf8; return res; }   // lib.rs:40 to_str()
    try {
        Cursor c{text};
        MaterialMap           materials;
        std::vector<Sphere>   spheres;
        std::vector<Triangle> triangles;

        c.skip_comment();                                                  // :342
        if (!statement_camera(c, res.camera)) throw Fail{ParseError::MissingCamera};
        c.skip_whitespace();
        c.skip_comment();                                                  // :353
        while (statement_material(c, materials, allow_emission)) { c.skip_whitespace(); c.skip_comment(); }
        while (statement_sphere(c, materials, spheres)) { c.skip_whitespace(); c.skip_comment(); }
        while (statement_triangle(c, materials, triangles)) { c.skip_whitespace(); c.skip_comment(); }
        if (!c.s.empty()) throw Fail{ParseError::WrongSyntax};             // :377-378
        res.world = World::make(std::move(spheres), std::move(triangles));
    } catch (const Fail& f) {
        res.error = f.e;
        res.world.reset();
    }
    return res;
}

// ------------------------------------------------------------------ image.rs

// image.rs:59-81 (ASCII P3): "P3\n{w} {h}\n255\n" then one "{r} {g} {b}\n" line per pixel, byte for byte what the
// reference's write!() calls produce.  The fast path: every channel value has a precomputed "ddd " token
// (256 x 4 bytes + its length), pixels are formatted straight into a 1 MiB buffer — no per-pixel printf.
bool write_image(const ColorU8* pixels, size_t width, size_t height, const char* path)
{
    FILE* f = path ? std::fopen(path, "w") : stdout;
    if (!f) return false;
    static const struct Tokens {
        char text[256][4]; unsigned char len[256];
        Tokens()
        {
            for (int v = 0; v < 256; ++v) {
                int n = 0;
                if (v >= 100) text[v][n++] = (char)('0' + v / 100);
                if (v >= 10) text[v][n++] = (char)('0' + (v / 10) % 10);
                text[v][n++] = (char)('0' + v % 10);
                len[v] = (unsigned char)n;
            }
        }
    } T;
    char head[64];
    const int hn = std::snprintf(head, sizeof head, "P3\n%zu %zu\n%d\n", width, height, 255);
    bool ok = std::fwrite(head, 1, (size_t)hn, f) == (size_t)hn;
    constexpr size_t kBuf = (size_t)1 << 20;
    std::vector<char> buf(kBuf + 16);
    size_t n = 0;
    auto put = [&](unsigned v, char sep) {
        const unsigned l = T.len[v];
        std::memcpy(&buf[n], T.text[v], 4);      // always 4 bytes: the tail is overwritten by what follows
        n += l;
        buf[n++] = sep;
    };
    for (size_t i = 0, count = width * height; i < count; ++i) {
        const ColorU8 c = pixels[i];
        put(c.r, ' '); put(c.g, ' '); put(c.b, '\n');
        if (n >= kBuf - 16) { ok = (std::fwrite(buf.data(), 1, n, f) == n) && ok; n = 0; }
    }
    if (n) ok = (std::fwrite(buf.data(), 1, n, f) == n) && ok;
    if (path) ok = (std::fclose(f) == 0) && ok;
    return ok;
}

// Binary PPM (P6): the same header with the magic P6, then 3 bytes per pixel.
bool write_image_p6(const ColorU8* pixels, size_t width, size_t height, const char* path)
{
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%zu %zu\n255\n", width, height);
    const size_t count = width * height;
    constexpr size_t kPix = (size_t)1 << 18;
    std::vector<unsigned char> rgb(kPix * 3);
    bool ok = true;
    for (size_t i0 = 0; i0 < count; i0 += kPix) {
        const size_t cnt = std::min(kPix, count - i0);
        for (size_t i = 0; i < cnt; ++i) {
            const ColorU8 c = pixels[i0 + i];
            rgb[3 * i + 0] = c.r; rgb[3 * i + 1] = c.g; rgb[3 * i + 2] = c.b;
        }
        ok = (std::fwrite(rgb.data(), 1, cnt * 3, f) == cnt * 3) && ok;
    }
    return (std::fclose(f) == 0) && ok;
}

bool write_image(const Framebuffer& fb, const char* path)
{
    return fb.pixels.size() == fb.width * fb.height && write_image(fb.pixels.data(), fb.width, fb.height, path);
}
bool write_image_p6(const Framebuffer& fb, const char* path)
{
    return fb.pixels.size() == fb.width * fb.height && write_image_p6(fb.pixels.data(), fb.width, fb.height, path);
}

uint32_t shard_tile_count(uint32_t height, uint32_t tile_rows, uint32_t index, uint32_t count)
{
    // full stripes of `count` tiles give every shard one tile; in the last, partial stripe a shard
    // has a tile when its (boustrophedon) position lies inside the remainder — rt_shard_tile
    const uint32_t tiles = (height + tile_rows - 1) / tile_rows;
    const uint32_t full  = tiles / count, rem = tiles % count;
    const uint32_t pos   = (full & 1u) ? count - 1u - index : index;
    return full + (pos < rem ? 1u : 0u);
}

}   // namespace rt
