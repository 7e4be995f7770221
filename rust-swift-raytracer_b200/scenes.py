"""Scene texts for the BASELINE.json configs, in the reference's world grammar
(parser.rs:326-335) so that they flow through load_world() unchanged.

* default_world(): the reference's default scene (raytracer/src/world.txt, SURVEY.md
  Appendix B): camera at the origin with aspect 1.77778, 9 materials, 8 spheres.
* example_world(): the same plus the two triangles of examples/c_raytracer.rs:42-43.
* synthetic_world(): the generated scenes of configs C3 / C5: a ground
  sphere plus small spheres on a jittered grid (and optional small triangles), mixed
  Diffuse / Metal / Dielectric.  Numbers are printed as fixed-point %.4f because
  parse_float (parser.rs:107-133) accepts no exponent and no '+'.

The generator's PRNG is xorshift32 on Python ints, so the text is identical on every
platform.
"""
from __future__ import annotations

ASPECT = "1.77778"

_DEFAULT_MATERIALS = [
    ("RED_DIFFUSE", "Diffuse color 1.0 0.0 0.0"),
    ("GREEN_DIFFUSE", "Diffuse color 0.0 1.0 0.0"),
    ("BLUE_DIFFUSE", "Diffuse color 0.0 0.0 1.0"),
    ("GROUND_MATERIAL", "Diffuse color 0.8 0.8 0.0"),
    ("BALL_MATERIAL", "Diffuse color 0.7 0.3 0.3"),
    ("METAL_MATERIAL_1", "Metal color 0.8 0.8 0.8 fuzz 0.3"),
    ("METAL_MATERIAL_2", "Metal color 0.8 0.6 0.2 fuzz 1.0"),
    ("MIRROR", "Metal color 0.9 0.9 0.9 fuzz 0.0"),
    ("GLASS", "Dielectric ir 1.5"),
]

# (centre, radius, material) in hit-test order
_DEFAULT_SPHERES = [
    ((0.0, -100.5, -1.0), 100.0, "GROUND_MATERIAL"),
    ((0.0, 0.0, -1.0), 0.5, "BALL_MATERIAL"),
    ((-1.0, 0.0, -1.0), 0.5, "METAL_MATERIAL_1"),
    ((1.0, 0.0, -1.0), 0.5, "GLASS"),
    ((0.0, 1.0, -2.0), 0.5, "MIRROR"),
    ((-3.0, 2.0, -3.0), 0.5, "RED_DIFFUSE"),
    ((0.0, 2.0, -3.0), 0.5, "GREEN_DIFFUSE"),
    ((3.0, 2.0, -3.0), 0.5, "BLUE_DIFFUSE"),
]


def default_world() -> str:
    lines = [f"camera origin 0.0 0.0 0.0 aspect {ASPECT};", ""]
    lines += [f"material {n} : {d};" for n, d in _DEFAULT_MATERIALS]
    lines.append("")
    for (x, y, z), r, m in _DEFAULT_SPHERES:
        lines.append(f"sphere center {x:.1f} {y:.1f} {z:.1f} radius {r:.1f} material {m};")
    return "\n".join(lines) + "\n"


def example_world() -> str:
    """examples/c_raytracer.rs:15-45: the default scene + two triangles at z = -0.5."""
    return default_world() + (
        "triangle v0 -0.1 -0.1 -0.5 v1 0.1 -0.1 -0.5 v2 -0.1 0.1 -0.5 material RED_DIFFUSE;\n"
        "triangle v0 -0.1 0.1 -0.5 v1 0.1 -0.1 -0.5 v2 0.1 0.1 -0.5 material GREEN_DIFFUSE;\n")


class _XorShift32:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFF or 1

    def u32(self) -> int:
        x = self.s
        x ^= (x << 13) & 0xFFFFFFFF
        x ^= x >> 17
        x ^= (x << 5) & 0xFFFFFFFF
        self.s = x
        return x

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        return lo + (hi - lo) * (self.u32() / 4294967296.0)


def _fmt(v: float) -> str:
    s = f"{v:.4f}"
    return "0.0000" if s == "-0.0000" else s


def synthetic_world(n_spheres: int = 1000, n_triangles: int = 0, seed: int = 1000) -> str:
    """Configs C3 (n_spheres=1000, seed=1000) and C5 (8000 spheres + 2000 triangles, seed=10000)."""
    rng = _XorShift32(seed)
    mats, spheres, tris = [], [], []

    def material(idx: int) -> str:
        name = f"M{idx}"
        k = rng.uniform()
        if k < 0.60:
            c = [rng.uniform(0.1, 0.9) for _ in range(3)]
            mats.append(f"material {name} : Diffuse color {_fmt(c[0])} {_fmt(c[1])} {_fmt(c[2])};")
        elif k < 0.85:
            c = [rng.uniform(0.1, 0.9) for _ in range(3)]
            f = rng.uniform(0.0, 0.5)
            mats.append(f"material {name} : Metal color {_fmt(c[0])} {_fmt(c[1])} {_fmt(c[2])} fuzz {_fmt(f)};")
        else:
            mats.append(f"material {name} : Dielectric ir 1.5000;")
        return name

    mats.append("material GROUND : Diffuse color 0.5000 0.5000 0.5000;")
    spheres.append("sphere center 0.0000 -1000.5000 -1.0000 radius 1000.0000 material GROUND;")

    n_small = max(n_spheres - 1, 0)
    # jittered grid with aspect ~ 37 x 27 (C3), scaled to hold n_small cells
    cols = max(1, int(round((n_small * 37.0 / 27.0) ** 0.5)))
    rows = max(1, (n_small + cols - 1) // cols)
    x0, x1, z0, z1 = -9.0, 9.0, -1.2, -14.0
    for i in range(n_small):
        cx, cz = i % cols, i // cols
        r = rng.uniform(0.05, 0.20)
        x = x0 + (x1 - x0) * (cx + 0.5 + rng.uniform(-0.3, 0.3)) / cols
        z = z0 + (z1 - z0) * (cz + 0.5 + rng.uniform(-0.3, 0.3)) / rows
        m = material(i)
        spheres.append(f"sphere center {_fmt(x)} {_fmt(r - 0.5)} {_fmt(z)} radius {_fmt(r)} material {m};")

    for j in range(n_triangles):
        cx = rng.uniform(x0, x1)
        cy = rng.uniform(-0.4, 1.5)
        cz = rng.uniform(z1, z0)
        v = [[c + rng.uniform(-0.08, 0.08) for c in (cx, cy, cz)] for _ in range(3)]
        m = material(n_small + j)
        tris.append("triangle " + " ".join(
            f"v{k} {_fmt(v[k][0])} {_fmt(v[k][1])} {_fmt(v[k][2])}" for k in range(3)) + f" material {m};")

    head = [f"camera origin 0.0 0.0 0.0 aspect {ASPECT};"]
    return "\n".join(head + mats + spheres + tris) + "\n"


def c3_world() -> str:
    return synthetic_world(1000, 0, seed=1000)


def c5_world() -> str:
    return synthetic_world(8000, 2000, seed=10000)
