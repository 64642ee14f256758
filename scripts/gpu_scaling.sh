#!/bin/bash
# Strong-scaling points of the headline bench on the GPUs this call was given (N = $1 ...), as the driver launches them;
# BIG=1 adds BASELINE config 5 (1024^3 x 1500 views) on the largest N.
set -u
TAG=${TAG:-r2}
mkdir -p gpurun_out
LAST=1
for N in "$@"; do
  LAST=$N
  if [ "$N" = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N"; fi
  echo "== $N GPU"; timeout 600 $L bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_512x720_${N}gpu.json 2> gpurun_out/${TAG}_bench_512x720_${N}gpu.err; echo "exit $?"
  tail -1 gpurun_out/${TAG}_bench_512x720_${N}gpu.json | cut -c1-200
done
if [ "${BIG:-0}" = 1 ]; then
  N=$LAST
  L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29599"
  echo "== 1024^3 x 1500 on $N GPU"; timeout 900 $L bench.py --gpus $N --size 1024 --views 1500 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_1024x1500_${N}gpu.json 2> gpurun_out/${TAG}_bench_1024x1500_${N}gpu.err; echo "exit $?"
  tail -1 gpurun_out/${TAG}_bench_1024x1500_${N}gpu.json | cut -c1-200
fi
