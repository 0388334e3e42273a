#!/bin/bash
# 2-GPU: DP parity test, then bench variants (SM reservation / NCCL channels)
mkdir -p gpurun_out
T=${1:-r2s}
timeout 400 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/${T}_dp.log 2>&1; echo "dp test exit $?"; grep -E "passed|failed|rel err|Error" gpurun_out/${T}_dp.log | tail -6
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu $EXTRA > gpurun_out/${T}_bench2_$name.json 2> gpurun_out/${T}_bench2_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${T}_bench2_$name.json'))
    dn=d.get('densenet18') or {}
    print('   resnet %.0f seq/s %.3f ms e2e %.0f | densenet %s | dp_parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], dn.get('value'), d.get('dp_parity')))
except Exception as e:
    print('   no json', e); import subprocess; print(open('gpurun_out/${T}_bench2_$name.err').read()[-1500:])
PY
}
EXTRA=""
run default A=1
EXTRA="--no-extra"
run reserve0 DEEPARDS_B200_DP_SM_RESERVE=0
run reserve4_ch4 DEEPARDS_B200_DP_SM_RESERVE=4 NCCL_MAX_NCHANNELS=4
run reserve8_ch8 DEEPARDS_B200_DP_SM_RESERVE=8 NCCL_MAX_NCHANNELS=8
run reserve16_ch16 DEEPARDS_B200_DP_SM_RESERVE=16 NCCL_MAX_NCHANNELS=16
run reserve0_ch2 DEEPARDS_B200_DP_SM_RESERVE=0 NCCL_MAX_NCHANNELS=2
EXTRA="--no-extra --scaling strong"
run strong A=1
