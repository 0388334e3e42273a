#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2w}
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu --no-extra > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${T}_$name.json')); print('$name: %.0f seq/s %.3f ms e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e:
    print('$name: no json', e); print(open('gpurun_out/${T}_$name.err').read()[-800:])
PY
}
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra > gpurun_out/${T}_1gpu.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/${T}_1gpu.json'));print('1 gpu: %.0f seq/s %.3f ms'%(d['value'],d['ms_per_step']))"
run real A=1
run noallreduce DEEPARDS_B200_DP_NO_ALLREDUCE=1
run noallreduce_reserve0 DEEPARDS_B200_DP_NO_ALLREDUCE=1 DEEPARDS_B200_DP_SM_RESERVE=0
run noallreduce_reserve0_whole DEEPARDS_B200_DP_NO_ALLREDUCE=1 DEEPARDS_B200_DP_SM_RESERVE=0 DEEPARDS_B200_DP_GRAPH=whole
run real_whole DEEPARDS_B200_DP_GRAPH=whole
run real_bigbuckets DEEPARDS_B200_DP_BUCKET_ELEMS=1048576
