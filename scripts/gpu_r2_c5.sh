#!/bin/bash
# tests, then C2 / C3 / C5 with the packed triangle stages; C5 with 1024 / 768 / 512-thread CTAs
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
line() { python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('$1', d['config']['workload'][:3], d['config']['kernel'][:5], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'], 'frac %.4f' % r['frac'])
"; }
B="--no-extras --no-cpu-baseline --no-e2e"
python bench.py --workload c2 --steps 40 $B 2>>gpurun_out/bench.err | line "c2"
python bench.py --workload c2 --steps 40 $B --fast-math 2>>gpurun_out/bench.err | line "c2"
python bench.py --workload c3 --steps 3 $B 2>>gpurun_out/bench.err | line "c3"
for v in "" b768 b512; do
  RT_LIB_VARIANT=$v python bench.py --workload c5 --steps 3 $B 2>>gpurun_out/bench.err | line "c5 [$v]"
  RT_LIB_VARIANT=$v python bench.py --workload c5 --steps 3 $B --fast-math 2>>gpurun_out/bench.err | line "c5 [$v]"
done
python bench.py --workload c3 --steps 3 $B --cull 2>>gpurun_out/bench.err | line "c3 cull"
python bench.py --workload c5 --steps 3 $B --cull 2>>gpurun_out/bench.err | line "c5 cull"
tail -5 gpurun_out/bench.err
