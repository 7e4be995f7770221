#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), then the ncu launch list and one full capture.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; cat gpurun_out/bench_c2.json
python bench.py --fast-math --no-cpu-baseline > gpurun_out/bench_c2_fast.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2_fast.json
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c3.json
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --fast-math > gpurun_out/bench_c3_fast.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c3_fast.json
python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c5.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rt_render_kernel -s 3 -c 1 -o gpurun_out/prof_c2_exact $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
CMD3="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --fast-math"
$CMD3 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rt_render_kernel -s 3 -c 1 -o gpurun_out/prof_c3_fast $CMD3 > gpurun_out/ncu_full3.log 2>&1
echo "ncu full c3 rc=$?"
