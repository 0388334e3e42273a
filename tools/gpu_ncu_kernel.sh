#!/bin/bash
# ncu --set full of one kernel family inside tools/kbench.py (source-level stall attribution)
mkdir -p gpurun_out
export KBENCH_NO_GRAPH=1 KBENCH_SHAPES=${SHAPES:-512x7}
WHAT=${WHAT:-convbn}; KREGEX=${KREGEX:-tc_conv_bn_kernel}; SKIP=${SKIP:-8}; TAG=${TAG:-a}
CMD="python tools/kbench.py $WHAT"
$CMD > gpurun_out/n_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/n_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c 2 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/n_ncu.log 2>&1
tail -2 gpurun_out/n_ncu.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
