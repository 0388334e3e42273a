#!/bin/bash
# tests (incl. the new rows) + wgrad A/B + config benches
TAG=${1:-x}
bash tools/gpu_tests.sh
echo "=== extra"; timeout 900 python -m pytest tests/test_extra_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/x_all.log 2>&1; echo "exit $?"; grep -E "passed|failed|gradcam:|worst" gpurun_out/x_all.log | tail -n 12
echo "=== kbench wgrad fused"; timeout 300 python tools/kbench.py conv > gpurun_out/kbench_${TAG}_fused.txt 2>&1; grep wgrad gpurun_out/kbench_${TAG}_fused.txt
echo "=== kbench wgrad 3 MMAs"; KBENCH_WGRAD_FUSE=0 timeout 300 python tools/kbench.py conv > gpurun_out/kbench_${TAG}_three.txt 2>&1; grep wgrad gpurun_out/kbench_${TAG}_three.txt
echo "=== configs"; timeout 900 python tools/bench_configs.py > gpurun_out/configs_${TAG}.jsonl 2> gpurun_out/configs_${TAG}.err; cat gpurun_out/configs_${TAG}.jsonl; tail -5 gpurun_out/configs_${TAG}.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; head -c 400 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
