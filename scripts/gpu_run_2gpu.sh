#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== 1 GPU 512"; timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_512c.json 2> gpurun_out/bench_1gpu_512c.err; echo "exit $?"; wc -l gpurun_out/bench_1gpu_512c.json
echo "== 2 GPU 512"; timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_512c.json 2> gpurun_out/bench_2gpu_512c.err; echo "exit $?"; wc -l gpurun_out/bench_2gpu_512c.json; grep -iE "rror|Traceback" gpurun_out/bench_2gpu_512c.err | head -5
