#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/tune_adjoint.py 512 180 fbg 2>&1 | tail -1
UNTILTED=1 timeout 300 python scripts/tune_adjoint.py 512 180 fbg 2>&1 | tail -1
