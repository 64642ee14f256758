#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=${1:-bench_512}
timeout 1200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/$OUT.json 2> gpurun_out/$OUT.err; echo "exit $?"; tail -1 gpurun_out/$OUT.json | cut -c1-300; tail -5 gpurun_out/$OUT.err
