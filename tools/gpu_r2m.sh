#!/bin/bash
# does the CTA-pair kernel survive under ncu?  (kbench conv at C=512: 6 fwd + 6 dgrad launches of tc_conv_pair_kernel)
mkdir -p gpurun_out
export KBENCH_NO_GRAPH=1 KBENCH_SHAPES=512x7
CMD="python tools/kbench.py conv"
$CMD > gpurun_out/m_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/m_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/m_launches.csv $CMD > gpurun_out/m_ncu.log 2>&1
echo "ncu exit $?"; tail -4 gpurun_out/m_launches.csv | cut -c1-220; tail -3 gpurun_out/m_ncu.log
