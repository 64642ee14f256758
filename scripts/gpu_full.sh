#!/bin/bash
# Round-end check: GPU test tier, smoke(), default bench line.
set -u
mkdir -p gpurun_out
TAG=${1:-r1_final}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "exit $?"; cat gpurun_out/bench_$TAG.json; tail -12 gpurun_out/bench_$TAG.err
