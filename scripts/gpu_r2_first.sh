#!/bin/bash
# Round-2 first call: GPU tests (new fused-pass / stealing / full-config tests included), the default
# bench line with its extra records, and an A/B of the C2 headline against the round-1 library (lib_base).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('C2 ms %.3f value %.0f e2e %.0f pinned %.0f frac %.4f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_pinned']['value'], d['roofline']['frac']))
for k,v in d.get('workloads',{}).items(): print(k, 'ms %.2f Mrays/s %.1f frac %.4f launches %d fused %d' % (v['ms_per_step'], v['value'], v['roofline']['frac'], v['launches_per_step'], v['passes_fused']))
print(d.get('ppm')); print(d['clocks']); print(d['cpu_baseline']['value'])
PY
for v in base ""; do
  RT_LIB_VARIANT=$v python bench.py --workload c2 --steps 30 --no-extras --no-cpu-baseline 2>>gpurun_out/bench.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('variant=[$v] C2 ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_frame']))"
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>>gpurun_out/bench.err; echo "ref rc=$?"
