#!/bin/bash
for MODE in "9=0" "" "9=0" ""; do
  DEEPARDS_B200_TC_DEBUG=$MODE timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_l2_x.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_l2_x.json')); r=d['roofline']['breakdown_ms_per_step']; print('[$MODE]', round(d['value']), round(d['ms_per_step'],4), d['final_loss'], r['dards_gbn_fwd'], r['dards_conv1d_fwd:tcgen05'], r['dards_conv1d_dgrad:tcgen05'], r['dards_gbn_bwd'])"
done
DEEPARDS_B200_TC_DEBUG="" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --backbone densenet18 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dense hint', round(d['value']), d['ms_per_step'])"
timeout 800 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_r1y.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/pytest_gpu_r1y.log
