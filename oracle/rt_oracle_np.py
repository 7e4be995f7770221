"""Second, independent restatement of the reference render loop in numpy float32 scalars.

TEST INFRASTRUCTURE ONLY (same rules as rt_oracle.c).  Purpose: the reference cannot be
compiled here, so the C oracle is cross-checked against a restatement written separately,
straight from the Rust sources, in a different language and style.  Pure-Python loops:
use on tiny frames only.  Every operation is a numpy.float32 scalar operation, i.e. a
correctly rounded IEEE binary32 op, never promoted to double.

Citations: /root/reference/raytracer/src/<file>:<line>.
"""
from __future__ import annotations

import numpy as np

F = np.float32
U32 = 0xFFFFFFFF
INF = F(np.inf)

ZERO, ONE, TWO, HALF = F(0.0), F(1.0), F(2.0), F(0.5)
T_MIN = F(0.001)
EPS = F(1e-8)
TWO32 = F(4294967296.0)          # `u32::MAX as f32`, random.rs:16


# ---------------------------------------------------------------- maths.rs
def vadd(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])            # :148
def vsub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])            # :154
def vscale(a, s): return (a[0] * s, a[1] * s, a[2] * s)                   # :204-209
def vdiv(a, s): return (a[0] / s, a[1] / s, a[2] / s)                     # :211-216
def vneg(a): return (-a[0], -a[1], -a[2])                                 # :218-220
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]             # :82


def nvec(x, y, z):                                                        # NVec3::new :111-118
    with np.errstate(all="ignore"):
        length = np.sqrt(x * x + y * y + z * z)
        return (x / length, y / length, z / length)


def cross(a, b):                                                          # :88-94
    return (a[1] * b[2] - a[2] * b[1], -(a[0] * b[2] - a[2] * b[0]), a[0] * b[1] - a[1] * b[0])


def near_zero(a):                                                         # :46-49
    return abs(a[0]) < EPS and abs(a[1]) < EPS and abs(a[2]) < EPS


def reflect(v, n):                                                        # :26-28
    return vsub(v, vscale(n, TWO * dot(v, n)))


def refract(uv, n, ratio):                                                # :31-36
    cos_theta = dot(vneg(uv), n)
    perp = vscale(vadd(uv, vscale(n, cos_theta)), ratio)
    with np.errstate(all="ignore"):
        par = vscale(n, -np.sqrt(abs(ONE - dot(perp, perp))))
    return vadd(perp, par)


# --------------------------------------------------------------- random.rs
class Rng:
    def __init__(self, seed):
        self.state = int(seed) & U32

    def next_u32(self):                                                   # :22-30
        x = self.state
        x ^= (x << 13) & U32
        x ^= x >> 17
        x ^= (x << 5) & U32
        self.state = x
        return x

    def f32(self):                                                        # :15-17
        return F(self.next_u32()) / TWO32       # np.float32(int) rounds to nearest even

    def bilateral(self):                                                  # :19-21
        return self.f32() * TWO - ONE


def _mix32(x):
    x ^= x >> 16
    x = (x * 0x7FEB352D) & U32
    x ^= x >> 15
    x = (x * 0x846CA68B) & U32
    x ^= x >> 16
    return x


def sample_seed(seed, pixel, sample):
    """Per-(pixel, sample) stream seed of the parallel path (DESIGN.md "RNG")."""
    h = _mix32((pixel ^ seed) & U32)
    h = _mix32((h + sample * 0x9E3779B9 + 0x85EBCA6B) & U32)
    return h if h else 0x9E3779B9


def random_unit_sphere(rng):                                              # common.rs:32-38
    x = rng.bilateral()
    y = rng.bilateral()
    z = rng.bilateral()
    return nvec(x, y, z)


# --------------------------------------------------------------- camera.rs
def camera_new_at(origin, aspect):                                        # :21-33
    o = tuple(F(c) for c in origin)
    vh = TWO
    vw = F(aspect) * vh
    return {"origin": o, "horizontal": (vw, ZERO, ZERO), "vertical": (ZERO, vh, ZERO),
            "llc": vsub(o, (vw / TWO, vh / TWO, ONE))}


def cast_ray(cam, s, t):                                                  # :84-89
    p = vsub(vadd(vadd(cam["llc"], vscale(cam["horizontal"], s)), vscale(cam["vertical"], t)), cam["origin"])
    return cam["origin"], nvec(*p)


# --------------------------------------------------------------- common.rs
DIFFUSE, METAL, DIELECTRIC, EMISSION = 0, 1, 2, 3


def sphere_hit(sph, o, d, t_max):                                         # :60-98
    center, radius, _ = sph
    oc = vsub(o, center)
    a = ONE                                                               # maths.rs:127
    half_b = dot(oc, d)
    c = dot(oc, oc) - radius * radius
    disc = half_b * half_b - a * c
    if disc < ZERO:
        return None
    with np.errstate(all="ignore"):
        sq = np.sqrt(disc)
    roots = [(-half_b - sq) / a, (-half_b + sq) / a]
    ok = [x for x in roots if T_MIN < x and x < t_max]                    # :88-92
    if not ok:
        return None
    t = min(ok)
    pos = vadd(o, vscale(d, t))                                           # Ray::at :20
    nrm = nvec(*vdiv(vsub(pos, center), radius))                          # :95
    return t, pos, nrm


def triangle_intersect(tri, o, d, t_max):                                 # :124-166
    v0, v1, v2, _normal, _ = tri
    n = cross(vsub(v1, v0), vsub(v2, v0))
    den = dot(n, d)
    if -EPS < den and den < EPS:
        return None
    dd = dot(n, v0)
    with np.errstate(all="ignore"):
        t = (dot(n, o) + dd) / den                                        # sic
    if t < T_MIN or t > t_max:
        return None
    p = vadd(o, vscale(d, t))
    for a, b in ((v0, v1), (v1, v2), (v2, v0)):
        if dot(n, cross(vsub(b, a), vsub(p, a))) < ZERO:
            return None
    return t, p


def world_hit(world, o, d):                                               # :237-258
    spheres, triangles = world
    closest, rec = INF, None
    for s in spheres:
        h = sphere_hit(s, o, d, closest)
        if h is not None:
            closest = h[0]
            rec = (h[0], h[1], h[2], s[2])
    best, mesh_rec = INF, None                                            # Mesh::hit :178-223
    for tr in triangles:
        h = triangle_intersect(tr, o, d, closest)
        if h is not None and h[0] < best:
            best = h[0]
            mesh_rec = (h[0], h[1], tr[3], tr[4])
    if mesh_rec is not None:
        rec = mesh_rec
    return rec


def scatter(mat, o, d, hit, rng):                                         # materials.rs:31-102
    t, pos, n, _ = hit
    kind, color, param = mat
    if kind == DIFFUSE:
        s = vadd(n, random_unit_sphere(rng))
        return color, (pos, n if near_zero(s) else nvec(*s))
    if kind == METAL:
        refl = reflect(d, n)
        direction = vadd(refl, vscale(random_unit_sphere(rng), param))
        if dot(direction, n) >= ZERO:
            return color, (pos, nvec(*direction))
        return color, None
    if kind == DIELECTRIC:
        if dot(d, n) >= ZERO:
            normal, ratio = vneg(n), ONE / param
        else:
            normal, ratio = n, param
        return (ONE, ONE, ONE), (pos, nvec(*refract(d, normal, ratio)))
    return color, None                                                    # Emission


def ray_color(o, d, world, rng, depth, counter):                          # common.rs:263-285
    final = (ONE, ONE, ONE)
    for _ in range(depth):
        counter[0] += 1
        hit = world_hit(world, o, d)
        if hit is not None:
            color, nxt = scatter(hit[3], o, d, hit, rng)
            final = (final[0] * color[0], final[1] * color[1], final[2] * color[2])
            if nxt is None:
                return final
            o, d = nxt
        else:
            t = HALF * (nvec(*d)[1] + ONE)
            sky = vadd(vscale((ONE, ONE, ONE), ONE - t), vscale((F(0.5), F(0.7), F(1.0)), t))
            return (final[0] * sky[0], final[1] * sky[1], final[2] * sky[2])
    return (ZERO, ZERO, ZERO)


def as_u8(x):
    if not (x == x) or x <= 0:
        return 0
    if x >= 255:
        return 255
    return int(x)


def make_world(spheres, triangles=()):
    """spheres: (center, radius, (kind, color, param)); triangles: (v0, v1, v2, (kind, color, param))."""
    def f3(v): return tuple(F(c) for c in v)
    def mat(m): return (m[0], f3(m[1]), F(m[2]))
    S = [(f3(c), F(r), mat(m)) for c, r, m in spheres]
    T = []
    for v0, v1, v2, m in triangles:
        v0, v1, v2 = f3(v0), f3(v1), f3(v2)
        nrm = nvec(*cross(vsub(v1, v0), vsub(v2, v0)))                    # Triangle::new :116-123
        T.append((v0, v1, v2, nrm, mat(m)))
    return (S, T)


def ray_trace(world, cam, width, height, spp, depth, *, serial=False, seed=2547549, fixed_jitter=False):
    """common.rs:320-361.  Returns (pixels[H,W,4] uint8, ray segments)."""
    out = np.zeros((height, width, 4), dtype=np.uint8)
    counter = [0]
    rng = Rng(seed)
    with np.errstate(all="ignore"):
        for row in range(height):
            for col in range(width):
                r = g = b = ZERO
                a = ONE                                                   # Color::new(0,0,0): alpha 1
                for s in range(spp):
                    if not serial:
                        rng = Rng(sample_seed(seed, row * width + col, s))
                    ju = HALF if fixed_jitter else rng.f32()
                    u = (F(col) + ju) / F(width - 1)
                    jv = HALF if fixed_jitter else rng.f32()
                    v = (F(row) + jv) / F(height - 1)
                    o, d = cast_ray(cam, u, v)
                    c = ray_color(o, d, world, rng, depth, counter)
                    r, g, b, a = r + c[0], g + c[1], b + c[2], a + ONE
                k = ONE / F(spp)
                px = (as_u8(np.sqrt(r * k) * F(255.999)), as_u8(np.sqrt(g * k) * F(255.999)),
                      as_u8(np.sqrt(b * k) * F(255.999)), as_u8(a * k * F(255.999)))
                out[height - row - 1, col] = px
    return out, counter[0]
