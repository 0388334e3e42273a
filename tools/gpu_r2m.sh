#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k "conv_bn" -q -p no:cacheprovider 2>&1 | tail -3
echo "--- convbn"; timeout 300 python tools/kbench.py convbn 2>&1 | tail -12
