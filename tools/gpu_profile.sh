#!/bin/bash
# ncu passes of one bench step (B200_PROFILING.md recipe).  Each ncu run follows a plain run of the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,gpu__time_duration.sum \
    --clock-control none -k regex:"tc_conv_kernel|tc_wgrad_kernel|gbn_|stem_" -s 220 -c 110 -o gpurun_out/prof_sections $CMD > gpurun_out/ncu_sections.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tc_conv_kernel" -s 60 -c 3 -o gpurun_out/prof_full_conv $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
