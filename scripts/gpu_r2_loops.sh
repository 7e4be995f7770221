#!/bin/bash
# Same-box A/B of the intersection-loop restructuring (lib_base = the library before it): full GPU test suite on the new
# library, then C2 / C3 / C5 exact (and C3 fast-math) on both.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
line() { python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print('$1', d['config']['workload'][:3], d['config']['kernel'][:5], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'], 'frac %.4f' % r['frac'])
"; }
B="--no-extras --no-cpu-baseline --no-e2e"
for v in base ""; do
  RT_LIB_VARIANT=$v timeout 120 python bench.py --workload c2 --steps 40 $B 2>>gpurun_out/bench.err | line "[$v]"
  RT_LIB_VARIANT=$v timeout 120 python bench.py --workload c3 --steps 3 $B 2>>gpurun_out/bench.err | line "[$v]"
  RT_LIB_VARIANT=$v timeout 120 python bench.py --workload c5 --steps 3 $B 2>>gpurun_out/bench.err | line "[$v]"
done
RT_LIB_VARIANT=base timeout 120 python bench.py --workload c2 --steps 40 $B 2>>gpurun_out/bench.err | line "[base]"
timeout 120 python bench.py --workload c2 --steps 40 $B 2>>gpurun_out/bench.err | line "[]"
timeout 120 python bench.py --workload c3 --steps 3 --fast-math $B 2>>gpurun_out/bench.err | line "[]"
tail -3 gpurun_out/bench.err
