#!/bin/bash
mkdir -p gpurun_out
export KBENCH_SHAPES=256x14
KBENCH_NO_GRAPH=1 timeout 300 python tools/kbench.py convbn > gpurun_out/r2o_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_bn -s 8 -c 2 -f -o gpurun_out/r2o_prof env KBENCH_NO_GRAPH=1 python tools/kbench.py convbn > gpurun_out/r2o_ncu.log 2>&1
ls -la gpurun_out/r2o_prof.ncu-rep; tail -2 gpurun_out/r2o_ncu.log
