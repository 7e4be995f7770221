#!/bin/bash
# Light ncu capture (a few sections, few replay passes) of the render kernel at FULL workload size.
set -u
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  NAME=$1; ARGS=$2; shift 2
  CMD="python bench.py $ARGS --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/plain_$NAME.log 2>&1 &&
  ncu --section SpeedOfLight --section WarpStateStats --section ComputeWorkloadAnalysis --section SchedulerStats --section Occupancy --section LaunchStats --section InstructionStats --clock-control none -k regex:rt_render_kernel -s 3 -c 1 -f -o gpurun_out/light_$NAME $CMD > gpurun_out/ncul_$NAME.log 2>&1
  echo "$NAME ncu rc=$?"
done
