#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2v}
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "wgrad" 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_model_parity_gpu.py tests/test_trainer_gpu.py tests/test_extra_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
for M in deterministic accumulate; do
DEEPARDS_B200_WGRAD=$M timeout 600 python bench.py --no-cpu --no-extra > gpurun_out/${T}_bench_$M.json 2> gpurun_out/${T}_bench_$M.err; echo "wgrad $M:"; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_$M.json'));print(d['value'], d['ms_per_step'], d['gpu_launches']); print({k:v for k,v in d['roofline']['breakdown_ms_per_step'].items() if 'wgrad' in k or 'unpack' in k or 'memset' in k})"
DEEPARDS_B200_WGRAD=$M timeout 600 python bench.py --no-cpu --no-extra --backbone densenet18 > gpurun_out/${T}_bench_dense_$M.json 2> gpurun_out/${T}_bench_dense_$M.err; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_dense_$M.json'));print('densenet', d['value'], d['ms_per_step'])"
done
