#!/bin/bash
# Round-end single-GPU measurements: default bench line (with the CPU baseline), the reference arm, BASELINE config 1-2
# size, and the opt-in z-quad kernels beside the default ones.  TAG names the outputs.
set -u
TAG=${TAG:-r2}
mkdir -p gpurun_out
echo "== bench 512x720"; timeout 900 python bench.py > gpurun_out/${TAG}_bench_512x720_1gpu.json 2> gpurun_out/${TAG}_bench_512x720_1gpu.err; echo "exit $?"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "exit $?"
echo "== bench 256x360"; timeout 600 python bench.py --size 256 --views 360 --no-cpu-baseline > gpurun_out/${TAG}_bench_256x360_1gpu.json 2> gpurun_out/${TAG}_bench_256x360_1gpu.err; echo "exit $?"
echo "== zquad vs default (512^3, 180 views)"
: > gpurun_out/${TAG}_zquad_vs_default.jsonl
timeout 300 python scripts/tune_adjoint.py 512 180 fg >> gpurun_out/${TAG}_zquad_vs_default.jsonl 2>&1
ZQUAD=1 timeout 300 python scripts/tune_adjoint.py 512 180 fg >> gpurun_out/${TAG}_zquad_vs_default.jsonl 2>&1
cat gpurun_out/${TAG}_zquad_vs_default.jsonl
for f in gpurun_out/${TAG}_bench_512x720_1gpu.json gpurun_out/${TAG}_bench_256x360_1gpu.json gpurun_out/${TAG}_bench_reference_arm.json; do tail -1 $f | cut -c1-300; done
