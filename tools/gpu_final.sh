#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke(), the bench (both arms), + the ncu launch list
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest -m gpu exit $?"; tail -n 3 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; head -c 200 gpurun_out/bench_ref_${TAG}.json; echo
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; head -c 300 gpurun_out/bench_${TAG}.json; echo; tail -3 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --backbone densenet18 --no-cpu > gpurun_out/bench_${TAG}_dense.json 2> gpurun_out/bench_${TAG}_dense.err; head -c 300 gpurun_out/bench_${TAG}_dense.json; echo
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv | head -12
