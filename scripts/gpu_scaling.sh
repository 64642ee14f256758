#!/bin/bash
# Strong-scaling points of the headline bench on the GPUs this call was given (N = $1 ...), as the driver launches them.
set -u
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N"; fi
  echo "== $N GPU"; timeout 600 $L bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_512_r2.json 2> gpurun_out/bench_${N}gpu_512_r2.err; echo "exit $?"
  tail -1 gpurun_out/bench_${N}gpu_512_r2.json | cut -c1-260
done
