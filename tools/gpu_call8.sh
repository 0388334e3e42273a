#!/bin/bash
for MODE in "9=0" "" "9=0" ""; do
  DEEPARDS_B200_TC_DEBUG=$MODE timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_l2_x.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_l2_x.json')); r=d['roofline']['breakdown_ms_per_step']; print('[$MODE]', round(d['value']), round(d['ms_per_step'],4), d['final_loss'], r['dards_gbn_bwd'], r['dards_conv1d_wgrad:tcgen05'], r['dards_conv1d_dgrad:tcgen05'])"
done
DEEPARDS_B200_TC_DEBUG="" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --backbone densenet18 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dense hint', round(d['value']), d['ms_per_step'])"
DEEPARDS_B200_TC_DEBUG="9=0" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --backbone densenet18 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dense plain', round(d['value']), d['ms_per_step'])"
