#!/bin/bash
mkdir -p gpurun_out
true
: > gpurun_out/tune.log
for L in tune/*.so; do TOMO_B200_LIB=$PWD/$L timeout 300 python scripts/tune_adjoint.py 512 180 fg >> gpurun_out/tune.log 2>&1; done
cat gpurun_out/tune.log
