#!/bin/bash
python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py -m gpu -x -q -k "conv_vs_oracle or vgg or resnet or single_pass" 2>&1 | tail -3
bash tools/gpu_call7.sh 2>&1 | head -45
