"""Generates tests/golden/frames.npz from the CPU oracle (oracle/rt_oracle.c).

    python tests/golden/make_golden.py

The reference ships no golden image and cannot be built here (no Rust toolchain), so these
vectors pin the ORACLE (per-sample RNG mode, the mode the GPU path is compared in, plus the
reference's serial mode) against drift; the oracle itself is pinned by the reference's own
known-answer tests and the independent numpy restatement (tests/test_oracle_*.py).
"""
import hashlib
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")]
import cases            # noqa: E402
import oracle_binding as ob   # noqa: E402

scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")


def main():
    out = {}
    for name, key, camera, W, H, spp, depth, fixed in cases.SMALL_CASES:
        cam, world = cases.oracle_scene(ob, scenes, key, camera)
        px, rays, _ = ob.ray_trace(world, cam, W, H, spp, depth, rng_mode=ob.RNG_PER_SAMPLE, fixed_jitter=fixed)
        out[name] = px
        out[name + "__rays"] = np.array([rays], dtype=np.uint64)
        print(f"{name:14s} {W}x{H} spp={spp} depth={depth} rays={rays} sha256={hashlib.sha256(px.tobytes()).hexdigest()[:16]}")
    # the reference's own serial-stream mode on a small frame (what `cargo run` would compute)
    cam, world = cases.oracle_scene(ob, scenes, "default", cases.LOOK_AT_CLI)
    px, rays, _ = ob.ray_trace(world, cam, 100, 56, 4, 8, rng_mode=ob.RNG_SERIAL)
    out["serial_small"] = px
    out["serial_small__rays"] = np.array([rays], dtype=np.uint64)
    np.savez_compressed(Path(__file__).parent / "frames.npz", **out)


if __name__ == "__main__":
    main()
