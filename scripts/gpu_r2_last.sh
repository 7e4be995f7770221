#!/bin/bash
# Last evidence call of round 2: GPU suite + smoke on the current library, same-box A/B of the small-scene kernel
# against the previous commit's library (lib_base), and the default bench line.  RT_LAST_BASE_LINE=1 also takes the
# default line of lib_base.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
line() { python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print('$1', d['config']['workload'][:3], d['config']['kernel'][:5], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'], 'frac %.4f' % r['frac'])
"; }
B="--no-extras --no-cpu-baseline --no-e2e"
for rep in 1 2; do
  RT_LIB_VARIANT=base timeout 120 python bench.py --workload c2 --steps 40 $B 2>>gpurun_out/bench.err | line "[base]"
  timeout 120 python bench.py --workload c2 --steps 40 $B 2>>gpurun_out/bench.err | line "[new ]"
done
timeout 200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
if [ "${RT_LAST_BASE_LINE:-0}" = 1 ]; then
  RT_LIB_VARIANT=base timeout 200 python bench.py > gpurun_out/bench_default_base.json 2>> gpurun_out/bench_default.err; echo "bench(base) rc=$?"
fi
python - <<'PY'
import json
for f in ("bench_default.json", "bench_default_base.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        w = d.get("workloads", {})
        print(f, "ms", d["ms_per_step"], "e2e", d["e2e"].get("ms_per_frame"), "frac", d["roofline"]["frac"],
              {k: (v.get("ms_per_step"), (v.get("roofline") or {}).get("frac")) for k, v in w.items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/bench_default.err
