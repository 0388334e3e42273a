#!/bin/bash
# bench (both backbones) + ncu launch list; $1 = tag
TAG=${1:-x}
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 1500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --steps 20 --warmup 5 --backbone densenet18 --no-cpu > gpurun_out/bench_${TAG}_dense.json 2> gpurun_out/bench_${TAG}_dense.err; tail -c 1200 gpurun_out/bench_${TAG}_dense.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
