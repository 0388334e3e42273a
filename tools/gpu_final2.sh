#!/bin/bash
TAG=${1:-x}
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest -m gpu exit $?"; tail -n 2 gpurun_out/pytest_gpu_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; head -c 260 gpurun_out/bench_${TAG}.json; echo; tail -2 gpurun_out/bench_${TAG}.err
