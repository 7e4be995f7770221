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
