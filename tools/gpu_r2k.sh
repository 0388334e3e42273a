#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k "conv_bn or tcgen05" -q -p no:cacheprovider 2>&1 | tail -3
export KBENCH_SHAPES=512x7,256x14
echo "--- plain conv: shared loop (2 x 160 columns)"; timeout 300 python tools/kbench.py conv 2>&1 | grep -v wgrad
echo "--- plain conv: shared loop, one 256-column tile"; KBENCH_CB_WIDE=1 timeout 300 python tools/kbench.py conv 2>&1 | grep -v wgrad
echo "--- convbn"; KBENCH_SHARED=0 timeout 300 python tools/kbench.py convbn 2>&1 | tail -4
echo "--- convbn per-tap loads"; KBENCH_SHARED=0 KBENCH_CB_PERTAP=1 timeout 300 python tools/kbench.py convbn 2>&1 | tail -4
