#!/bin/bash
# N-GPU call (gpurun --gpus N): the multi-GPU tests, then the bench's C4 line with and without work stealing.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests -m gpu -q -k "gpus or per_gpu or stealing or unbalanced or another_device or two_devices or many_gpus" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_multi.log
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 5 --warmup 3 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_c4_n$N.err
run --steps 5 --warmup 3 --no-steal --no-extras > gpurun_out/bench_c4_n${N}_nosteal.json 2>> gpurun_out/bench_c4_n$N.err; echo "bench(no steal) rc=$?"
run --steps 5 --warmup 3 --no-row-gather --no-extras > gpurun_out/bench_c4_n${N}_norowgather.json 2>> gpurun_out/bench_c4_n$N.err; echo "bench(no row gather) rc=$?"
run --steps 20 --warmup 3 --workload c2 --no-extras > gpurun_out/bench_c2_n${N}.json 2>> gpurun_out/bench_c4_n$N.err; echo "bench(c2) rc=$?"
run --steps 20 --warmup 3 --workload c2 --no-row-gather --no-extras > gpurun_out/bench_c2_n${N}_norowgather.json 2>> gpurun_out/bench_c4_n$N.err; echo "bench(c2, no row gather) rc=$?"
python - <<PY
import json
for f in ("gpurun_out/bench_c4_n$N.json", "gpurun_out/bench_c4_n${N}_nosteal.json", "gpurun_out/bench_c4_n${N}_norowgather.json", "gpurun_out/bench_c2_n${N}.json", "gpurun_out/bench_c2_n${N}_norowgather.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split('/')[-1], 'ms %.3f value %.0f e2e %.0f' % (d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value', 0)),
          'stolen', d['config'].get('stolen_slots_first_step'), d.get('strong_scaling'), d.get('parity'), d.get('e2e_one_process'))
PY
