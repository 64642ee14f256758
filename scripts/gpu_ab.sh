#!/bin/bash
# A/B: parity tests of the default library, then adjoint timing of every tune/*.so beside it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "adjoint or back or determin or drop_in" 2>&1 | tail -4
: > gpurun_out/tune.log
for L in tomography_alignment_b200/libtomo_b200.so tune/*.so; do TOMO_B200_LIB=$PWD/$L timeout 300 python scripts/tune_adjoint.py 512 180 ${WHICH:-b} >> gpurun_out/tune.log 2>&1; done
cat gpurun_out/tune.log
