#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --size 256 --views 48 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'ray_kernel|adjoint_tile' -s 3 -c 3 -o gpurun_out/prof_r1c $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
