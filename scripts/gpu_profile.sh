#!/bin/bash
# Round-end ncu evidence at the benchmark size: launch list of one bench run (host-built phantom keeps torch's phantom
# kernels out of it), then one --set full capture per hot kernel.  TAG names the output files (default r2).
set -u
TAG=${TAG:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --host-phantom"
timeout 900 $CMD > gpurun_out/plain_profile.log 2> gpurun_out/plain_profile.err || { echo "plain run failed"; tail -5 gpurun_out/plain_profile.err; exit 1; }
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_512x720.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
for K in ${KERNELS:-ray_kernel_forward ray_kernel_gradient adjoint_tile_kernel}; do
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_$K $CMD > gpurun_out/ncu_full_$K.log 2>&1
  echo "ncu $K exit $?"
done
ls -la gpurun_out/*.ncu-rep | tail -7
