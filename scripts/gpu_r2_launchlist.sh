#!/bin/bash
# ncu launch list (gpu__time_duration only, no clock control) of the final library's C2 bench command — the same command
# exited 0 without ncu in scripts/gpu_r2_last.sh.  Per-launch times are cold-cache and serialised: compare shares, not absolutes.
set -u
mkdir -p gpurun_out
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2_final.csv \
  python bench.py --workload c2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_final.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/launches_c2_final.csv | cut -c1-200
