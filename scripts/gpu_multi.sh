#!/bin/bash
# Multi-GPU check: N = $1 ranks (torchrun, one process per GPU), C4 strong scaling + the 2-device test.
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu or many_gpus or per_gpu" > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; cat gpurun_out/bench_c4_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_c4_n$n.json 2> gpurun_out/bench_c4_n$n.err
    echo "rc=$?"; cat gpurun_out/bench_c4_n$n.json; tail -5 gpurun_out/bench_c4_n$n.err
  fi
done
