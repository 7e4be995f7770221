#!/bin/bash
# A/B of the staged edge-stage records (RtSceneView::hot_parts): full GPU suite on the default, the triangle / fuzz tests
# with each forced combination of parts, then C5 with the round's previous staging (RT_HOT_PARTS=4: r*r only) against the default.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for m in 0 1 3; do
  RT_HOT_PARTS=$m timeout 200 python -m pytest tests -m gpu -x -q -k "fuzz or c5 or tris or c3" > gpurun_out/pytest_parts$m.log 2>&1; echo "parts=$m rc=$?"; tail -1 gpurun_out/pytest_parts$m.log
done
line() { python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print('$1', d['config']['workload'][:3], d['config']['kernel'][:5], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'], 'frac %.4f' % r['frac'], 'smem', d['config'].get('launch',{}).get('smem_bytes'))
"; }
B="--no-extras --no-cpu-baseline --no-e2e"
for m in 4 "" 4 "" 1; do
  RT_HOT_PARTS=$m timeout 120 python bench.py --workload c5 --steps 3 $B 2>>gpurun_out/bench.err | line "[parts=$m]"
done
RT_HOT_PARTS=4 timeout 120 python bench.py --workload c5 --steps 3 --fast-math $B 2>>gpurun_out/bench.err | line "[parts=4]"
timeout 120 python bench.py --workload c5 --steps 3 --fast-math $B 2>>gpurun_out/bench.err | line "[parts=]"
tail -3 gpurun_out/bench.err
