#!/bin/bash
# FFMA2 (packed) sphere filter: tests, then C3/C5 with one and two paths per lane, both kernels.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
line() { python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; e=d.get('e2e') or {}
    print('$1', d['config']['workload'][:3], d['config']['kernel'][:5], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'], 'frac %.4f' % r['frac'])
"; }
B="--no-extras --no-cpu-baseline --no-e2e"
for np in 1 2; do
  for fm in "" "--fast-math"; do
    RT_PATHS_PER_LANE=$np python bench.py --workload c3 --steps 3 $B $fm 2>>gpurun_out/bench.err | line "np=$np"
    RT_PATHS_PER_LANE=$np python bench.py --workload c5 --steps 3 $B $fm 2>>gpurun_out/bench.err | line "np=$np"
  done
done
python bench.py --workload c2 --steps 30 $B 2>>gpurun_out/bench.err | line "c2"
tail -5 gpurun_out/bench.err
