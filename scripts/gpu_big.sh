#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== 1024^3 x 1500 (BASELINE config 5), one step, device-resident"
timeout 1500 python bench.py --size 1024 --views 1500 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; echo "exit $?"; tail -1 gpurun_out/bench_1024.json | cut -c1-400; tail -6 gpurun_out/bench_1024.err
nvidia-smi --query-gpu=memory.used --format=csv
