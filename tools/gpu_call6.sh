#!/bin/bash
TAG=${1:-x}
bash tools/gpu_tests.sh
grep -E "FAILED|Error" gpurun_out/k_simt.log gpurun_out/k_tc.log gpurun_out/m_all.log | head -10
timeout 300 python tools/kbench.py conv 2>&1 | grep wgrad
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; head -c 300 gpurun_out/bench_${TAG}.json; echo; tail -3 gpurun_out/bench_${TAG}.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"tc_wgrad_kernel|stem_bwd" -s 40 -c 20 -f -o /tmp/prof_wg_${TAG} $CMD > gpurun_out/ncu_wg_${TAG}.log 2>&1
ncu -i /tmp/prof_wg_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_wg_${TAG}_raw.csv 2> /dev/null
python tools/ncu_summary.py gpurun_out/prof_wg_${TAG}_raw.csv | head -8
