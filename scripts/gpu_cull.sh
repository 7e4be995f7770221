#!/bin/bash
# Cull-mode check: the cull tests + C3/C5 benches with and without --cull.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cull or fuzz" > gpurun_out/pytest_cull.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_cull.log
run() { python bench.py "$@" --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print(d['config']['workload'][:3], d['config']['kernel'][:5], d['config']['sphere_walk'][:10], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'])
"; }
run --workload c3 --steps 2 --warmup 3 --cull
run --workload c3 --steps 2 --warmup 3 --cull --fast-math
run --workload c5 --steps 2 --warmup 3 --cull
run --workload c5 --steps 2 --warmup 3 --cull --fast-math
