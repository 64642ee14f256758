#!/bin/bash
# DRAM traffic of each hot kernel at the real benchmark size: one ncu --set full capture per kernel.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 $CMD > gpurun_out/plain_traffic.log 2> gpurun_out/plain_traffic.err || exit 1
for K in adjoint_tile "ray_kernel<0>" ; do
  N=$(echo $K | tr -cd 'a-z0-9_')
  timeout 1500 ncu --set full --clock-control none -k regex:"$K" -s 1 -c 1 -o gpurun_out/prof_r1_512_$N $CMD > gpurun_out/ncu_traffic_$N.log 2>&1
  echo "ncu $K exit $?"
done
