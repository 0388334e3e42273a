#!/bin/bash
mkdir -p gpurun_out
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-extra --no-cpu --steps 60 --warmup 10 > gpurun_out/q_${label}.json 2> gpurun_out/q_${label}.err || { echo "$label FAILED"; tail -3 gpurun_out/q_${label}.err; return; }
  python - <<PY
import json
d=json.load(open('gpurun_out/q_${label}.json'))
bd=d['roofline']['breakdown_ms_per_step']
print('%-14s %9.0f seq/s  %.4f ms | dgrad %.4f fwd %.4f convbn %.4f wgrad %.4f' % ('${label}', d['value'], d['ms_per_step'], bd.get('dards_conv1d_dgrad:tcgen05',0), bd.get('dards_conv1d_fwd:tcgen05',0), bd.get('dards_conv1d_bn_fwd',0), bd.get('dards_conv1d_wgrad_accum',0)))
PY
}
run base A=1
run pair DEEPARDS_B200_TC_DEBUG=17=1
run base2 A=1
run pair2 DEEPARDS_B200_TC_DEBUG=17=1
