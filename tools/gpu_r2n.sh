#!/bin/bash
# ncu --set full of the fused conv+BN kernel alone (C=512 L=7 and C=256 L=14), with source-level stall attribution
mkdir -p gpurun_out
export KBENCH_NO_GRAPH=1 KBENCH_SHAPES=${SHAPES:-512x7}
CMD="python tools/kbench.py convbn"
$CMD > gpurun_out/n_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/n_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_bn_kernel" -s 8 -c 2 -f -o gpurun_out/convbn_${TAG:-a} $CMD > gpurun_out/n_ncu.log 2>&1
tail -3 gpurun_out/n_ncu.log
ls -la gpurun_out/convbn_${TAG:-a}.ncu-rep
