"""Shared parity cases: (name, scene text builder, camera spec, W, H, spp, depth, fixed_jitter)."""
import math

import numpy as np

VFOV_90 = float(np.float32(math.pi) / np.float32(2.0))   # Radians(PI / 2.0), main.rs:87
LOOK_AT_CLI = ((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), VFOV_90, 1.77778)   # main.rs:86-88
LOOK_AT_TILTED = ((1.5, 1.0, 1.0), (0.0, 0.25, -1.0), (0.0, 1.0, 0.0), 0.9, 1.5)

# name, scene, camera ("file" = camera of the scene text, or a look_at tuple), W, H, spp, depth, fixed
SMALL_CASES = [
    ("c1_det",        "default", LOOK_AT_CLI,    400, 224, 1, 8, True),    # BASELINE config 1, deterministic mode
    ("c1_8spp",       "default", LOOK_AT_CLI,    400, 224, 8, 8, False),
    ("abi_16spp",     "default", "file",         160, 90, 16, 8, False),   # what render() does (lib.rs:51)
    ("example_tris",  "example", "file",         200, 200, 4, 8, False),   # examples/c_raytracer.rs
    ("tilted_cam",    "example", LOOK_AT_TILTED, 96, 64, 4, 8, False),
    ("odd_size",      "example", "file",         123, 77, 3, 5, False),
    ("depth1",        "default", "file",         64, 36, 4, 1, False),
    ("c3_small",      "c3",      "file",         96, 54, 2, 8, False),
    ("c5_small",      "c5mini",  "file",         64, 36, 2, 16, False),
]


def scene_text(scenes, key):
    if key == "default":
        return scenes.default_world()
    if key == "example":
        return scenes.example_world()
    if key == "c3":
        return scenes.c3_world()
    if key == "c5mini":
        return scenes.synthetic_world(800, 200, seed=10000)
    if key == "c5":
        return scenes.c5_world()
    raise KeyError(key)


def oracle_scene(ob, scenes, key, camera):
    cam, world = ob.parse_input(scene_text(scenes, key))
    if camera != "file":
        cam = ob.camera_new_look_at(*camera)
    return cam, world


def product_scene(rt, scenes, key, camera):
    h = rt.load_world(scene_text(scenes, key))
    if camera != "file":
        h.set_camera_look_at(*camera)
    return h


# ---- random worlds for fuzzing the conservative filters (spheres / triangle plane + edge stages) ----

def random_world(seed: int, n_spheres: int = 70, n_triangles: int = 12) -> str:
    """A world text with awkward geometry: radii over six decades, far-away and huge primitives,
    spheres containing the camera, needle-thin / degenerate / far-from-origin triangles.
    >= 64 spheres so that the FILTER kernels run."""
    import numpy as np
    rng = np.random.default_rng(seed)

    def f(v):
        s = f"{float(v):.6f}"
        return "0.000000" if s == "-0.000000" else s

    lines = [f"camera origin {f(rng.uniform(-1, 1))} {f(rng.uniform(-0.5, 1))} {f(rng.uniform(-1, 1))} aspect 1.5;"]
    mats = []
    for i in range(8):
        k = rng.integers(0, 3)
        c = rng.uniform(0.05, 1.0, 3)
        if k == 0:
            mats.append(f"material M{i} : Diffuse color {f(c[0])} {f(c[1])} {f(c[2])};")
        elif k == 1:
            mats.append(f"material M{i} : Metal color {f(c[0])} {f(c[1])} {f(c[2])} fuzz {f(rng.choice([0.0, 0.1, 0.5, 1.0]))};")
        else:
            mats.append(f"material M{i} : Dielectric ir {f(rng.choice([1.0, 1.33, 1.5, 2.4]))};")
    lines += mats
    for i in range(n_spheres):
        kind = rng.integers(0, 10)
        if kind == 0:      # huge, far below (ground-like)
            r = 10.0 ** rng.uniform(1, 4)
            c = (rng.uniform(-5, 5), -r - rng.uniform(0.3, 1.0), rng.uniform(-5, 5))
        elif kind == 1:    # tiny
            r = 10.0 ** rng.uniform(-3, -1.5)
            c = rng.uniform(-2, 2, 3) + np.array([0, 0, -3.0])
        elif kind == 2:    # far away
            r = rng.uniform(5, 50)
            c = rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(2, 3.5)
        elif kind == 3:    # contains the camera
            r = rng.uniform(3, 30)
            c = rng.uniform(-1, 1, 3)
        else:
            r = rng.uniform(0.05, 0.8)
            c = rng.uniform(-4, 4, 3) + np.array([0, 0, -5.0])
        lines.append(f"sphere center {f(c[0])} {f(c[1])} {f(c[2])} radius {f(r)} material M{rng.integers(0, 8)};")
    for j in range(n_triangles):
        kind = rng.integers(0, 6)
        base = rng.uniform(-3, 3, 3) + np.array([0, 0, -4.0])
        if kind == 0:      # needle: two vertices almost coincide
            v = [base, base + rng.uniform(-2, 2, 3), None]
            v[2] = v[1] + rng.uniform(-1, 1, 3) * 1e-4
        elif kind == 1:    # degenerate: collinear
            d = rng.uniform(-1, 1, 3)
            v = [base, base + d, base + 2.0 * d]
        elif kind == 2:    # large
            v = [base + rng.uniform(-30, 30, 3) for _ in range(3)]
        elif kind == 3:    # far from the origin
            far = base * 10.0 ** rng.uniform(1.5, 3)
            v = [far + rng.uniform(-1, 1, 3) for _ in range(3)]
        else:
            v = [base + rng.uniform(-0.7, 0.7, 3) for _ in range(3)]
        lines.append("triangle " + " ".join(f"v{k} {f(v[k][0])} {f(v[k][1])} {f(v[k][2])}" for k in range(3)) +
                     f" material M{rng.integers(0, 8)};")
    return "\n".join(lines) + "\n"
