#!/bin/bash
# N=2: distributed GPU tests, then the two weak-scaling lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -5
bash tools/gpu_call_nN.sh 2 r02x
