#!/bin/bash
# ncu passes of one bench step (B200_PROFILING.md recipe).  Each ncu run follows a plain run of the same command.
# $1 = tag.  Outputs: launch list (all kernels), --set full capture of the dominant kernel classes.
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 417 -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gbn_bwd_cached|gbn_fwd_cached|tc_conv_kernel|tc_wgrad_kernel" -s 240 -c 100 -o gpurun_out/prof_full_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep
