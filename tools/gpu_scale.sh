#!/bin/bash
# N-GPU weak-scaling bench, launched the way the driver does; $1 = N, $2 = tag
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
for BB in resnet18 densenet18; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 --backbone $BB > gpurun_out/bench_${N}gpu_${BB}_${TAG}.json 2> gpurun_out/bench_${N}gpu_${BB}_${TAG}.err
  echo "rc=$?"; head -c 600 gpurun_out/bench_${N}gpu_${BB}_${TAG}.json; echo; tail -2 gpurun_out/bench_${N}gpu_${BB}_${TAG}.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${N}gpu_ref_${TAG}.json 2>/dev/null; head -c 300 gpurun_out/bench_${N}gpu_ref_${TAG}.json
