#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2i}
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k "conv_bn or tcgen05" -q -p no:cacheprovider 2>&1 | tail -4
echo "--- plain conv: shared loop on (C>128)"; timeout 300 python tools/kbench.py conv 2>&1 | grep -v wgrad
echo "--- plain conv: shared loop everywhere"; KBENCH_SHARED=2 timeout 300 python tools/kbench.py conv 2>&1 | grep -v wgrad
echo "--- convbn (two-kernel side uses old conv)"; KBENCH_SHARED=0 timeout 300 python tools/kbench.py convbn 2>&1 | tail -12
