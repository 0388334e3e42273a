#!/bin/bash
# A/B sweep of scheduling knobs on the full step (ResNet-18 and DenseNet-18), one bench line each.
mkdir -p gpurun_out
run() {  # label, env assignments...
  local label=$1; shift
  for bb in resnet18 densenet18; do
    env "$@" timeout 300 python bench.py --no-extra --no-cpu --backbone $bb --steps 60 --warmup 10 > gpurun_out/x_${label}_${bb}.json 2> gpurun_out/x_${label}_${bb}.err || { echo "$label $bb FAILED"; tail -3 gpurun_out/x_${label}_${bb}.err; continue; }
    python - <<PY
import json
d=json.load(open('gpurun_out/x_${label}_${bb}.json'))
print('%-28s %-10s %9.0f seq/s  %.4f ms  launches/step %d' % ('${label}', '${bb}', d['value'], d['ms_per_step'], d['gpu_launches']//d['steps']))
PY
  done
}
DEEPARDS_B200_TC_DEBUG="15=1,16=1" timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "gbn or bn" 2>&1 | tail -2
run base A=1
run bn_k DEEPARDS_B200_TC_DEBUG=16=1
run bn_shift_k DEEPARDS_B200_TC_DEBUG=15=1,16=1
run bn_shift DEEPARDS_B200_TC_DEBUG=15=1
run tile_balance DEEPARDS_B200_TC_DEBUG=8=1
run fuse3 DEEPARDS_B200_FUSE_BN=3
run base2 A=1
