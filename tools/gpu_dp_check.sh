#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/dp_${TAG}.log 2>&1; echo "dp test exit $?"; grep -E "passed|failed|rel err|Error" gpurun_out/dp_${TAG}.log | tail -8
for BB in resnet18 densenet18; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 --backbone $BB --no-cpu > gpurun_out/bench_2gpu_${BB}_${TAG}.json 2> gpurun_out/bench_2gpu_${BB}_${TAG}.err
  echo "$BB rc=$?"; head -c 250 gpurun_out/bench_2gpu_${BB}_${TAG}.json; echo
done
