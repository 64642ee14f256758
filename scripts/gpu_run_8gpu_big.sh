#!/bin/bash
# 8 GPUs: the headline 512^3 x 720 bench, then BASELINE config 5 (1024^3 x 1500, volume replicated, views sharded), device-resident.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
echo "== 8 GPU 512"; timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu_512_v9.json 2> gpurun_out/bench_8gpu_512_v9.err; echo "exit $?"; tail -1 gpurun_out/bench_8gpu_512_v9.json | cut -c1-300
echo "== 8 GPU 1024x1500"; timeout 900 $TR bench.py --gpus 8 --size 1024 --views 1500 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_8gpu_1024.json 2> gpurun_out/bench_8gpu_1024.err; echo "exit $?"; tail -1 gpurun_out/bench_8gpu_1024.json | cut -c1-300; grep -E "rror|Traceback" gpurun_out/bench_8gpu_1024.err | tail -5
