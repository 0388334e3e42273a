#!/bin/bash
TAG=${1:-x}
bash tools/gpu_tests.sh
echo "=== extra"; timeout 900 python -m pytest tests/test_extra_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/x_all.log 2>&1; echo "exit $?"; grep -E "passed|failed|gradcam:|worst" gpurun_out/x_all.log | tail -n 8
grep -E "FAILED|Error" gpurun_out/k_tc.log | head -10
echo "=== kbench balanced"; timeout 300 python tools/kbench.py conv > gpurun_out/kbench_${TAG}_bal.txt 2>&1; grep -E "fwd|dgrad" gpurun_out/kbench_${TAG}_bal.txt
echo "=== kbench widest"; KBENCH_TILE_BALANCE=0 timeout 300 python tools/kbench.py conv > gpurun_out/kbench_${TAG}_wide.txt 2>&1; grep -E "fwd|dgrad" gpurun_out/kbench_${TAG}_wide.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; head -c 300 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
