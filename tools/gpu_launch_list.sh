#!/bin/bash
# ncu launch list of one bench step (the --metrics gpu__time_duration.sum pass of B200_PROFILING.md), after a plain run
TAG=${1:-r02b}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-extra"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/launches_${TAG}.csv | cut -c1-200
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}_summary.txt && head -30 gpurun_out/launches_${TAG}_summary.txt
