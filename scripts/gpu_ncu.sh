#!/bin/bash
# ncu full captures of the render kernel: gpu_ncu.sh name1 "bench args 1" [name2 "bench args 2" ...]
set -u
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  NAME=$1; ARGS=$2; shift 2
  CMD="python bench.py $ARGS --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
  $CMD > gpurun_out/plain_$NAME.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:rt_render_kernel -s 3 -c 1 -f -o gpurun_out/prof_$NAME $CMD > gpurun_out/ncu_$NAME.log 2>&1
  echo "$NAME ncu rc=$?"; tail -1 gpurun_out/plain_$NAME.log | cut -c1-200
done
