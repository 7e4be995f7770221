#!/bin/bash
# Quick iteration call: GPU tests + short benches of the main workloads (+ optional ncu of one).
# usage: gpu_quick.sh [ncu-workload] [ncu-extra-flags]
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
run() { python bench.py "$@" --no-cpu-baseline 2>>gpurun_out/bench.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; e=d.get('e2e') or {}
    print(d['config']['workload'][:3], d['config']['kernel'][:5], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'], 'e2e %.1f' % e.get('value',0), 'frac %.4f' % r['frac'], 'peak %.1f' % r['peak'], d['clocks'])
"; }
run --steps 20 --warmup 3
run --steps 20 --warmup 3 --fast-math
run --workload c3 --steps 2 --warmup 3 --no-e2e
run --workload c3 --steps 2 --warmup 3 --no-e2e --fast-math
run --workload c5 --steps 2 --warmup 3 --no-e2e
run --workload c5 --steps 2 --warmup 3 --no-e2e --fast-math
run --workload c3 --steps 2 --warmup 3 --no-e2e --cull
run --workload c3 --steps 2 --warmup 3 --no-e2e --cull --fast-math
run --workload c5 --steps 2 --warmup 3 --no-e2e --cull
if [ $# -ge 1 ]; then
  W=$1; shift
  CMD="python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-e2e $*"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:rt_render_kernel -s 3 -c 1 -f -o gpurun_out/prof_quick $CMD > gpurun_out/ncu_quick.log 2>&1
  echo "ncu rc=$?"
fi
