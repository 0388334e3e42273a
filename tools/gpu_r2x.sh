#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2x}; N=${2:-2}
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --no-extra > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${T}_$name.json')); print('$name: %.0f seq/s %.3f ms e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e:
    print('$name: no json', e); print(open('gpurun_out/${T}_$name.err').read()[-800:])
PY
}
run b2M_r8 DEEPARDS_B200_DP_BUCKET_ELEMS=2097152
run b2M_r0 DEEPARDS_B200_DP_BUCKET_ELEMS=2097152 DEEPARDS_B200_DP_SM_RESERVE=0
run b2M_r4 DEEPARDS_B200_DP_BUCKET_ELEMS=2097152 DEEPARDS_B200_DP_SM_RESERVE=4 NCCL_MAX_NCHANNELS=4
run b4M_r0 DEEPARDS_B200_DP_BUCKET_ELEMS=4194304 DEEPARDS_B200_DP_SM_RESERVE=0
run b4M_r0_ch32 DEEPARDS_B200_DP_BUCKET_ELEMS=4194304 DEEPARDS_B200_DP_SM_RESERVE=0 NCCL_MAX_NCHANNELS=32
run b1M_r8 DEEPARDS_B200_DP_BUCKET_ELEMS=1048576
