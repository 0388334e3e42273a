#!/bin/bash
# ncu launch list of one bench step (the --metrics gpu__time_duration.sum pass of B200_PROFILING.md), after a plain run
TAG=${1:-r02b}
mkdir -p gpurun_out
# The bench process dies with SIGSEGV (exit 11, no CUDA error) as soon as ncu PROFILES a tc_conv_pair_kernel launch of the
# training step (the same kernel launched from tools/kbench.py profiles fine, and launches inside ncu's skipped range are
# harmless; compute-sanitizer is closed on this pool, so this was not chased further).  The step's launch list is therefore
# taken with the single-CTA kernels in place of the pair kernels (debug keys 17 = 0, 19 = 0); the pair kernel's own durations under
# ncu are in profiles/r02_pair_kernel_ncu.txt.
export DEEPARDS_B200_TC_DEBUG=${DEEPARDS_B200_TC_DEBUG:-17=0,19=0}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-extra"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/launches_${TAG}.csv | cut -c1-200
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}_summary.txt && head -30 gpurun_out/launches_${TAG}_summary.txt
