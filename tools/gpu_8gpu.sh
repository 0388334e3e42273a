#!/bin/bash
# 8-GPU: scaling bench variants + DenseNet + GradCAM over the sharded recording
mkdir -p gpurun_out
N=${1:-8}; T=${2:-r2t}
run() { name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu $EXTRA > gpurun_out/${T}_bench${N}_$name.json 2> gpurun_out/${T}_bench${N}_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${T}_bench${N}_$name.json'))
    dn=d.get('densenet18') or {}
    print('   resnet %.0f seq/s %.3f ms e2e %.0f | densenet %s %s | dp_parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], dn.get('value'), dn.get('ms_per_step'), d.get('dp_parity')))
except Exception as e:
    print('   no json', e); print(open('gpurun_out/${T}_bench${N}_$name.err').read()[-1200:])
PY
}
EXTRA=""
run default A=1
EXTRA="--no-extra"
run reserve16_ch16 DEEPARDS_B200_DP_SM_RESERVE=16 NCCL_MAX_NCHANNELS=16
run reserve0 DEEPARDS_B200_DP_SM_RESERVE=0
run whole DEEPARDS_B200_DP_GRAPH=whole
EXTRA="--no-extra --scaling strong"
run strong A=1
EXTRA="--no-extra --backbone densenet18 --scaling strong"
run strong_dense A=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/bench_gradcam_dp.py > gpurun_out/${T}_gradcam${N}.json 2> gpurun_out/${T}_gradcam${N}.err; echo "gradcam rc=$?"; cat gpurun_out/${T}_gradcam${N}.json; tail -2 gpurun_out/${T}_gradcam${N}.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu --backbone densenet18 --no-extra > gpurun_out/${T}_bench4_dense.json 2> gpurun_out/${T}_bench4_dense.err; echo "4-GPU densenet rc=$?"; head -c 200 gpurun_out/${T}_bench4_dense.json; echo
