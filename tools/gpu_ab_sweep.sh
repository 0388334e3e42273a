#!/bin/bash
# A/B sweep of kernel-variant switches on the full step, one bench line each:  tools/gpu_ab_sweep.sh "label:ENV=VALUE ..." ...
# e.g. tools/gpu_ab_sweep.sh base: share:DEEPARDS_B200_TC_DEBUG=14=1 nopair:DEEPARDS_B200_TC_DEBUG=17=0
mkdir -p gpurun_out
BACKBONES=${BACKBONES:-resnet18}
for spec in "$@"; do
  label=${spec%%:*}; assign=${spec#*:}
  for bb in $BACKBONES; do
    env ${assign:-A=1} timeout 300 python bench.py --no-extra --no-cpu --backbone $bb --steps 60 --warmup 10 > gpurun_out/x_${label}_${bb}.json 2> gpurun_out/x_${label}_${bb}.err || { echo "$label $bb FAILED"; tail -3 gpurun_out/x_${label}_${bb}.err; continue; }
    python - <<PY
import json
d=json.load(open('gpurun_out/x_${label}_${bb}.json'))
print('%-22s %-10s %9.0f seq/s  %.4f ms  launches/step %d' % ('${label}', '${bb}', d['value'], d['ms_per_step'], d['gpu_launches']//d['steps']))
PY
  done
done
