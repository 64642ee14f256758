#!/bin/bash
# 8 GPUs: BASELINE config 5 (1024^3 x 1500, volume replicated, views sharded), device-resident.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
echo "== 8 GPU 1024x1500"; timeout 900 $TR bench.py --gpus 8 --size 1024 --views 1500 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_8gpu_1024_v11.json 2> gpurun_out/bench_8gpu_1024_v11.err; echo "exit $?"; tail -1 gpurun_out/bench_8gpu_1024_v11.json | cut -c1-300; grep -E "rror|Traceback" gpurun_out/bench_8gpu_1024_v11.err | tail -5
