#!/bin/bash
# adjoint: parity tests, then tilted and untilted timing of the default library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "adjoint or back or determin or drop_in or separable or sirt" 2>&1 | tail -3
timeout 300 python scripts/tune_adjoint.py 512 180 b
UNTILTED=1 timeout 300 python scripts/tune_adjoint.py 512 180 b
