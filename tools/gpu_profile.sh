#!/bin/bash
# ncu passes of one bench step (B200_PROFILING.md recipe).  Each ncu run follows a plain run of the same command.
# $1 = tag, $2 = backbone.  Outputs: launch list (all kernels), --set full capture of one step's heavy kernels exported
# as raw CSV (the .ncu-rep is kept only when it is small enough to travel).
TAG=${1:-x}
BB=${2:-resnet18}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-extra --backbone $BB"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}_summary.txt; head -28 gpurun_out/launches_${TAG}_summary.txt
K="gbn_bwd|gbn_fwd|gbn_apply|tc_conv|tc_wgrad|unpack_wgrad|stem_|dropout"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 300 -c 150 -f -o /tmp/prof_full_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i /tmp/prof_full_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_full_${TAG}_raw.csv 2> /dev/null
SZ=$(stat -c %s /tmp/prof_full_${TAG}.ncu-rep)
echo "report bytes: $SZ"
if [ "$SZ" -lt 40000000 ]; then cp /tmp/prof_full_${TAG}.ncu-rep gpurun_out/; fi
python tools/ncu_summary.py gpurun_out/prof_full_${TAG}_raw.csv > gpurun_out/full_step_${TAG}_summary.txt; head -40 gpurun_out/full_step_${TAG}_summary.txt
