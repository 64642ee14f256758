#!/bin/bash
# ncu launch list (per-launch durations) of one bench run at the benchmark size; phantom generation included, so -c is generous.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --host-phantom"
timeout 900 $CMD > gpurun_out/plain_profile.log 2> gpurun_out/plain_profile.err || { echo "plain run failed"; exit 1; }
timeout 2000 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_512x720.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r1_launches_512x720.csv
