"""Quick GPU sanity run (development aid): CUDA path vs oracle on a few configs + timings."""
import importlib, sys, time, math, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np
rt = importlib.import_module("rust-swift-raytracer_b200")
scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")
import oracle_binding as ob

print("devices", rt.device_count(), flush=True)

def compare(name, text, W, H, spp, depth, fixed, look_at=False, fast=False):
    h = rt.load_world(text)
    cam, w = ob.parse_input(text)
    if look_at:
        vf = float(np.float32(math.pi) / np.float32(2.0))
        h.set_camera_look_at((0, 0, 0), (0, 0, -1), (0, 1, 0), vf, 1.77778)
        cam = ob.camera_new_look_at((0, 0, 0), (0, 0, -1), (0, 1, 0), vf, 1.77778)
    assert np.array_equal(h.camera_floats(), cam.floats()), (h.camera_floats(), cam.floats())
    fb = rt.Framebuffer(W, H)
    st = rt.RenderStats()
    t = time.time()
    rt.render_with_options(fb, h, rt.Options(spp, depth, fixed_jitter=fixed, fast_math=fast), st)
    dt = time.time() - t
    px, rays, _ = ob.ray_trace(w, cam, W, H, spp, depth, fixed_jitter=fixed)
    d = np.abs(fb.pixels.astype(int) - px.astype(int))
    rmse = float(np.sqrt((d[:, :, :3].astype(float) ** 2).mean()))
    print(f"{name:10s} {W}x{H} spp={spp} d={depth} fixed={int(fixed)} fast={int(fast)}: equal={np.array_equal(fb.pixels, px)} "
          f"maxdiff={d.max()} rmse={rmse:.4f} nz={(d.max(axis=2)>0).mean():.5f} rays gpu={st.rays} cpu={rays} kernel_ms={st.kernel_ms:.3f} "
          f"call_s={dt:.3f} grid={st.grid}", flush=True)
    return fb

dw, ew = scenes.default_world(), scenes.example_world()
compare("C1-det", dw, 400, 224, 1, 8, True, look_at=True)
compare("C1-50", dw, 400, 224, 50, 8, False, look_at=True)
compare("example", ew, 200, 200, 16, 8, False)
compare("odd-size", ew, 123, 77, 3, 5, False)
compare("C3-small", scenes.c3_world(), 192, 108, 2, 8, False)
compare("C5-small", scenes.synthetic_world(800, 200, seed=10000), 128, 72, 2, 16, False)
compare("C1-fast", dw, 400, 224, 50, 8, False, look_at=True, fast=True)
fb = compare("ex-fast", ew, 200, 200, 16, 8, False, fast=True)
from PIL import Image
out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
Image.fromarray(fb.pixels[:, :, :3]).save(out / "example_fast.png")

# timings: C2
h = rt.load_world(dw)
for fast in (False, True):
    fb = rt.Framebuffer(1920, 1080, pinned=True)
    st = rt.RenderStats()
    for i in range(4):
        t = time.time()
        rt.render_with_options(fb, h, rt.Options(64, 8, fast_math=fast), st)
        dt = time.time() - t
        print(f"C2 fast={int(fast)} iter{i}: kernel_ms={st.kernel_ms:.3f} call_ms={dt*1e3:.3f} rays={st.rays} Mrays/s={st.rays/st.kernel_ms/1e3:.1f}", flush=True)
    Image.fromarray(fb.pixels[:, :, :3]).save(out / f"c2_fast{int(fast)}.png")
print("fp32 peak TFLOP/s", rt.measure_fp32_peak(), flush=True)
h3 = rt.load_world(scenes.c3_world())
for fast in (False, True):
    fb = rt.Framebuffer(1920, 1080, pinned=True)
    st = rt.RenderStats()
    t = time.time()
    rt.render_with_options(fb, h3, rt.Options(4, 8, fast_math=fast), st)
    print(f"C3@4spp fast={int(fast)}: kernel_ms={st.kernel_ms:.3f} rays={st.rays} Mrays/s={st.rays/st.kernel_ms/1e3:.1f} smem={st.smem_bytes}", flush=True)
    Image.fromarray(fb.pixels[:, :, :3]).save(out / f"c3_fast{int(fast)}.png")
