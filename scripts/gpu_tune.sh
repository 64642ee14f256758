#!/bin/bash
# Time the operators selected by $WHICH (default: adjoint) for the default library and every tune/*.so.
mkdir -p gpurun_out
: > gpurun_out/tune.log
for L in tomography_alignment_b200/libtomo_b200.so tune/*.so; do TOMO_B200_LIB=$PWD/$L timeout 300 python scripts/tune_adjoint.py ${N:-512} ${VIEWS:-180} ${WHICH:-b} >> gpurun_out/tune.log 2>&1; done
cat gpurun_out/tune.log
