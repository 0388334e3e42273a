#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2y}; N=${2:-8}
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --no-extra $EXTRA > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${T}_$name.json')); print('$name: %.0f seq/s %.3f ms e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e:
    print('$name: no json', e); print(open('gpurun_out/${T}_$name.err').read()[-800:])
PY
}
EXTRA=""
run one_bucket_r0 DEEPARDS_B200_DP_BUCKET_ELEMS=4194304 DEEPARDS_B200_DP_SM_RESERVE=0
run one_bucket_r0_ch32 DEEPARDS_B200_DP_BUCKET_ELEMS=4194304 DEEPARDS_B200_DP_SM_RESERVE=0 DEEPARDS_B200_NCCL_CHANNELS=32
run two_buckets_r0 DEEPARDS_B200_DP_BUCKET_ELEMS=2097152 DEEPARDS_B200_DP_SM_RESERVE=0
run five_buckets_r8_ch8 DEEPARDS_B200_NCCL_CHANNELS=8
EXTRA="--backbone densenet18"
run dense_r0 DEEPARDS_B200_DP_SM_RESERVE=0
run dense_r8_ch8 DEEPARDS_B200_NCCL_CHANNELS=8
