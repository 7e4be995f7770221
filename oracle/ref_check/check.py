#!/usr/bin/env python3
"""Pin the oracle to the reference at IMAGE level (needs a Rust toolchain for the reference side).

    python oracle/ref_check/check.py                       # oracle only: prints / verifies the expected digests
    python oracle/ref_check/check.py --rust-ppm ref.ppm --samples 50 --depth 8 [--width 400]

The oracle's SERIAL mode consumes ONE xorshift32 stream (seed 2547549, random.rs:9) in the reference's
row -> column -> sample -> bounce order (common.rs:320-361), so its PPM must equal, byte for byte, the one
`cargo run` of oracle/ref_check (the unmodified crate) writes for the same world, camera, size, samples, depth.
expected.json holds the sha256 of the oracle's PPMs for BASELINE config 1 (world.txt, new_look_at camera,
400x224, depth 8) at 50 spp (what src/main.rs renders) and at 1 spp.

Caveat stated once: the camera uses f32::tan(PI/4) (camera.rs:53); glibc's tanf returns exactly 1.0 there (image
height 224).  A libm that returns 0.99999994 would shift the camera by one ulp and the digests with it.
"""
import argparse
import hashlib
import importlib
import json
import math
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import oracle_binding as ob  # noqa: E402


def oracle_ppm(samples: int, depth: int, width: int = 400, world_text: str | None = None) -> bytes:
    if world_text is None:
        world_text = importlib.import_module("rust-swift-raytracer_b200.scenes").default_world()
    _, world = ob.parse_input(world_text)
    vfov = float(np.float32(math.pi) / np.float32(2.0))                      # Radians(PI / 2.0), main.rs:87
    cam = ob.camera_new_look_at((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), vfov, 1.77778)
    aspect = np.float32(cam.floats()[6]) / np.float32(cam.floats()[10])       # horizontal.x / vertical.y, camera.rs:70-72
    height = int(np.float32(width) / aspect)                                  # main.rs:92
    px, _, _ = ob.ray_trace(world, cam, width, height, samples, depth, rng_mode=ob.RNG_SERIAL, threads=1)
    with tempfile.NamedTemporaryFile(suffix=".ppm") as f:
        ob.write_image(px, f.name)                                            # image.rs:59-81
        return Path(f.name).read_bytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rust-ppm")
    ap.add_argument("--samples", type=int, default=50)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--width", type=int, default=400)
    ap.add_argument("--world", help="world text file (default: the repository's copy of raytracer/src/world.txt)")
    ap.add_argument("--update", action="store_true", help="rewrite expected.json from the oracle")
    a = ap.parse_args()
    expected_path = HERE / "expected.json"
    if a.rust_ppm:
        text = Path(a.world).read_text() if a.world else None
        mine = oracle_ppm(a.samples, a.depth, a.width, text)
        theirs = Path(a.rust_ppm).read_bytes()
        if mine == theirs:
            print(f"IDENTICAL: {len(mine)} bytes, sha256 {hashlib.sha256(mine).hexdigest()} — the oracle reproduces the reference")
            return 0
        m, t = mine.split(b"\n"), theirs.split(b"\n")
        first = next((i for i, (x, y) in enumerate(zip(m, t)) if x != y), min(len(m), len(t)))
        print(f"DIFFERENT: oracle {len(mine)} bytes / reference {len(theirs)} bytes; first differing line {first + 1}: "
              f"oracle {m[first] if first < len(m) else None!r} reference {t[first] if first < len(t) else None!r}")
        return 1
    got = {f"c1_{s}spp_depth8_400x224_serial": hashlib.sha256(oracle_ppm(s, 8)).hexdigest() for s in (1, 50)}
    if a.update or not expected_path.exists():
        expected_path.write_text(json.dumps(got, indent=1) + "\n")
        print("wrote", expected_path)
    want = json.loads(expected_path.read_text())
    print(json.dumps(got, indent=1))
    return 0 if got == want else 1


if __name__ == "__main__":
    sys.exit(main())
