#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "voxel_driven" 2>&1 | tail -15
timeout 300 python scripts/tune_adjoint.py 512 180 v 2>&1 | tail -1
