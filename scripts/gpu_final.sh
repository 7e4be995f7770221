#!/bin/bash
# Round-end style run on 1 GPU: tests, smoke, both bench arms, launch list, interactive latency.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; cat gpurun_out/bench_reference.json
python - <<'PY'
import importlib, time, sys
sys.path.insert(0, ".")
rt = importlib.import_module("rust-swift-raytracer_b200"); scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")
h = rt.load_world(scenes.default_world())
for (W, H) in ((400, 224), (1920, 1080)):
    fb = rt.Framebuffer(W, H, pinned=True)
    rt.render(fb, h)
    for name, f in (("render() 16spp", lambda: rt.render(fb, h)),
                    ("progressive 1spp/call", lambda: rt.render_progressive(fb, h, rt.Options(1, 8)))):
        ts = []
        for i in range(30):
            rt.move_camera_position(h, 0.001, 0.0, 0.0)        # what a key press does (GameView.swift:198-216)
            t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
        ts.sort(); print(f"interactive {W}x{H} {name}: median {ts[15]*1e3:.3f} ms, p90 {ts[27]*1e3:.3f} ms")
PY
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
