#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
for d in 2 4 8; do
  if [ $d -le $N ]; then
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --devices $d 2>>gpurun_out/oneproc.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('C2 1-GPU e2e %.1f Mrays/s (%.3f ms)' % (d['e2e']['value'], d['e2e']['ms_per_frame']), '| one process:', json.dumps(d.get('e2e_one_process')))
"
  fi
done
python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --devices $N 2>>gpurun_out/oneproc.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('C3 1-GPU e2e %.1f Mrays/s (%.3f ms)' % (d['e2e']['value'], d['e2e']['ms_per_frame']), '| one process:', json.dumps(d.get('e2e_one_process')))
"
