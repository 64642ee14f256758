#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== $N GPU 512"; timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_512.json 2> gpurun_out/bench_${N}gpu_512.err; echo "exit $?"; tail -1 gpurun_out/bench_${N}gpu_512.json | cut -c1-600; grep -E "bench|rror" gpurun_out/bench_${N}gpu_512.err | tail -12
