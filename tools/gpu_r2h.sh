#!/bin/bash
mkdir -p gpurun_out
export KBENCH_SHAPES=512x7,256x14
for B in 4 6 8; do echo "--- plain conv, shared loop, b_stages=$B"; KBENCH_CB_BSTAGES=$B timeout 300 python tools/kbench.py conv 2>&1 | grep -v wgrad; done
echo "--- fused convbn b_stages=6"; KBENCH_SHARED=0 KBENCH_CB_BSTAGES=6 timeout 300 python tools/kbench.py convbn 2>&1 | tail -4
export KBENCH_SHAPES=512x7
KBENCH_NO_GRAPH=1 timeout 300 python tools/kbench.py conv > gpurun_out/r2h_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_bn -s 6 -c 2 -f -o gpurun_out/r2h_prof env KBENCH_NO_GRAPH=1 python tools/kbench.py conv > gpurun_out/r2h_ncu.log 2>&1
ls -la gpurun_out/r2h_prof.ncu-rep; tail -3 gpurun_out/r2h_ncu.log
