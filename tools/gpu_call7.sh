#!/bin/bash
for MODE in default dgrad_first default dgrad_first; do
  DEEPARDS_B200_BWD_ORDER=$MODE timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_order_$MODE.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_order_$MODE.json')); print('$MODE', round(d['value']), d['ms_per_step'], d['final_loss'])"
done
