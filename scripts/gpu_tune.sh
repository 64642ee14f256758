#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune.log
for L in tune/*.so; do UNTILTED=1 TOMO_B200_LIB=$PWD/$L timeout 300 python scripts/tune_adjoint.py 512 180 g >> gpurun_out/tune.log 2>&1; done
cat gpurun_out/tune.log
