#!/bin/bash
TAG=${1:-x}
bash tools/gpu_tests.sh
grep -E "FAILED|Error" gpurun_out/k_simt.log gpurun_out/k_tc.log gpurun_out/m_all.log | head -10
echo "=== extra"; timeout 600 python -m pytest tests/test_extra_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/x_all.log 2>&1; echo "exit $?"; grep -E "passed|failed" gpurun_out/x_all.log | tail -n 2
timeout 300 python tools/kbench.py stem 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; head -c 300 gpurun_out/bench_${TAG}.json; echo; tail -3 gpurun_out/bench_${TAG}.err
