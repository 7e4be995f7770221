"""Small renders covering every kernel variant, every kernel variant once (compute-sanitizer is closed on this pool, so this is a plain smoke)."""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rt = importlib.import_module("rust-swift-raytracer_b200")
scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")

def go(name, text, W, H, spp, depth, **kw):
    h = rt.load_world(text)
    fb = rt.Framebuffer(W, H)
    st = rt.RenderStats()
    rt.render_with_options(fb, h, rt.Options(spp, depth, **kw), st)
    print(name, W, H, spp, "rays", st.rays, "block", st.block, "filtered", st.filtered, "items", st.sample_items, "resident", st.resident, flush=True)

for fast in (False, True):
    go("default", scenes.default_world(), 67, 35, 3, 8, fast_math=fast)
    go("example+tris", scenes.example_world(), 41, 23, 2, 8, fast_math=fast)
    go("example items", scenes.example_world(), 41, 23, 3, 8, fast_math=fast, sample_items=True)
    go("shard 1/3", scenes.example_world(), 41, 23, 2, 8, fast_math=fast, shard_index=1, shard_count=3, tile_rows=4)
    go("c3 filter", scenes.c3_world(), 24, 14, 1, 4, fast_math=fast)
    go("c5 filter+tris 1024", scenes.c5_world(), 16, 9, 1, 3, fast_math=fast)
    go("global path", scenes.synthetic_world(15000, 500, seed=77), 8, 5, 1, 2, fast_math=fast)
    go("c3 cull", scenes.c3_world(), 24, 14, 1, 4, fast_math=fast, group_cull=True)
    go("c5 cull+tris", scenes.c5_world(), 16, 9, 1, 3, fast_math=fast, group_cull=True)
    go("global cull", scenes.synthetic_world(15000, 500, seed=77), 8, 5, 1, 2, fast_math=fast, group_cull=True)
go("empty", "camera origin 0.0 0.0 0.0 aspect 1.5;", 9, 5, 2, 3)
go("1x1", scenes.example_world(), 1, 1, 2, 3)
print("selftest", rt.selftest_division(1 << 16, 1))
print("done")
