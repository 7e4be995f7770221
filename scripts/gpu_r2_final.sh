#!/bin/bash
# Round-2 evidence run on 1 GPU: tests, smoke, both bench arms, launch list of the default bench, ncu --set full of the
# C2 / C3 / C5 render kernels (reduced spp: ncu replays the kernel ~40 times) and of the ordered-sum kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err; echo "ref rc=$?"
python bench.py --fast-math --steps 30 > gpurun_out/bench_c2_fast.json 2>> gpurun_out/bench_default.err
for w in c3 c5; do
  python bench.py --workload $w --steps 3 --fast-math --no-e2e --no-cpu-baseline > gpurun_out/bench_${w}_fast.json 2>> gpurun_out/bench_default.err
  python bench.py --workload $w --steps 3 --cull --no-e2e --no-cpu-baseline > gpurun_out/bench_${w}_cull.json 2>> gpurun_out/bench_default.err
done
python - <<'PY'
import importlib, time, sys
sys.path.insert(0, ".")
rt = importlib.import_module("rust-swift-raytracer_b200"); scenes = importlib.import_module("rust-swift-raytracer_b200.scenes")
h = rt.load_world(scenes.default_world())
for (W, H) in ((400, 224), (1920, 1080)):
    for pinned in (True, False):
        fb = rt.Framebuffer(W, H, pinned=pinned)
        rt.render(fb, h)
        for name, f in (("render() 16spp", lambda: rt.render(fb, h)),
                        ("progressive 1spp/call", lambda: rt.render_progressive(fb, h, rt.Options(1, 8)))):
            ts = []
            for i in range(30):
                rt.move_camera_position(h, 0.001, 0.0, 0.0)        # what a key press does (GameView.swift:198-216)
                t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
            ts.sort(); print(f"interactive {W}x{H} {'pinned' if pinned else 'pageable'} {name}: median {ts[15]*1e3:.3f} ms, p90 {ts[27]*1e3:.3f} ms")
PY
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
bash scripts/gpu_ncu.sh r02_c2 "--workload c2 --spp 16" r02_c3 "--workload c3 --spp 8" r02_c5 "--workload c5 --spp 2"
