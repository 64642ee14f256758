#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== 2 GPU 256"; timeout 900 $TR bench.py --gpus 2 --size 256 --views 360 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_256.json 2> gpurun_out/bench_2gpu_256.err; echo "exit $?"; tail -1 gpurun_out/bench_2gpu_256.json | cut -c1-1500; tail -3 gpurun_out/bench_2gpu_256.err
echo "== 1 GPU 512"; timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_512b.json 2> gpurun_out/bench_1gpu_512b.err; echo "exit $?"; tail -1 gpurun_out/bench_1gpu_512b.json | cut -c1-400
echo "== 2 GPU 512"; timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_512.json 2> gpurun_out/bench_2gpu_512.err; echo "exit $?"; tail -1 gpurun_out/bench_2gpu_512.json | cut -c1-1800; tail -3 gpurun_out/bench_2gpu_512.err
echo "== reference arm under torchrun"; timeout 600 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-size 128 > gpurun_out/bench_ref_2.json 2> gpurun_out/bench_ref_2.err; echo "exit $?"; tail -1 gpurun_out/bench_ref_2.json | cut -c1-600
