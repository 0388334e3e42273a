#!/bin/bash
mkdir -p gpurun_out
DEEPARDS_B200_TC_DEBUG=21=4 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "wgrad" 2>&1 | tail -3
for a in 0 2 4 8; do
  echo "== wgrad L2 prefetch distance = $a"
  KBENCH_WGRAD_PREFETCH=$a timeout 200 python tools/kbench.py conv 2>&1 | grep 'wgacc'
done
