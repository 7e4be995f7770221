#!/bin/bash
# C5: cull-record prefetch distance and edge-stage margin variants (built with build.py RT_BUILD_VARIANT)
set -u
mkdir -p gpurun_out
line() { python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('$1', d['config']['workload'][:3], d['config']['kernel'][:5], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'], 'frac %.4f' % r['frac'])
"; }
B="--no-extras --no-cpu-baseline --no-e2e"
for v in "" pf4 pf8 pf16 m125 pf8m125; do
  RT_LIB_VARIANT=$v python bench.py --workload c5 --steps 3 $B 2>>gpurun_out/bench.err | line "c5 [$v]"
done
RT_LIB_VARIANT=pf8m125 python -m pytest tests -m gpu -q -x -k "fuzz or c5 or tris or group_cull" 2>&1 | tail -3
tail -3 gpurun_out/bench.err
