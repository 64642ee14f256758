#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
echo "== bench 512"; timeout 1200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_512_r10.json 2> gpurun_out/bench_512_r10.err; echo "exit $?"; cat gpurun_out/bench_512_r10.json; tail -5 gpurun_out/bench_512_r10.err
