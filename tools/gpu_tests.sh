#!/bin/bash
# Runs the GPU test-suite in isolated processes (a faulting tcgen05 kernel must not take the fp32 results with it).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -q --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/$name.log; }
run k_simt tests/test_kernels_gpu.py -m gpu -k "not tcgen05"
run m_fp32 tests/test_model_parity_gpu.py -m gpu -k "not bf16"
run k_tc tests/test_kernels_gpu.py -m gpu -k "tcgen05"
run m_bf16 tests/test_model_parity_gpu.py -m gpu -k "bf16"
