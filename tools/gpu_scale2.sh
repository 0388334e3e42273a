#!/bin/bash
# 2-GPU A/B of the multi-rank step: whole-step graph with captured NCCL vs per-segment graphs; $1 = tag
TAG=${1:-x}
mkdir -p gpurun_out
for MODE in whole segments; do
 for BB in resnet18 densenet18; do
  DEEPARDS_B200_DP_GRAPH=$MODE timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 --backbone $BB --no-cpu > gpurun_out/bench_2gpu_${BB}_${MODE}_${TAG}.json 2> gpurun_out/bench_2gpu_${BB}_${MODE}_${TAG}.err
  echo "$MODE $BB rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_2gpu_${BB}_${MODE}_${TAG}.json"))
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["cuda_graph"], d["final_loss"])
except Exception as e:
    print("no json", e)
PY
  grep -i "warn\|error" gpurun_out/bench_2gpu_${BB}_${MODE}_${TAG}.err | tail -3
 done
done
timeout 600 python -m pytest tests/test_extra_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -4
