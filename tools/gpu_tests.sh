#!/bin/bash
# Runs the GPU test-suite in isolated processes (a faulting tcgen05 kernel must not take the fp32 results with it).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -q --tb=short -p no:cacheprovider -s > gpurun_out/$name.log 2>&1; echo "exit $?"; grep -E "passed|failed" gpurun_out/$name.log | tail -n 3; }
run k_simt tests/test_kernels_gpu.py -m gpu -k "not tcgen05"
run k_tc tests/test_kernels_gpu.py -m gpu -k "tcgen05"
run m_all tests/test_model_parity_gpu.py tests/test_trainer_gpu.py -m gpu
