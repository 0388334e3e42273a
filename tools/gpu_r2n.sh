#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2n}
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_parity_gpu.py tests/test_trainer_gpu.py tests/test_extra_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for F in 0 2 0 2; do
DEEPARDS_B200_FUSE_BN=$F timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_fuse$F.json 2> gpurun_out/${T}_bench_fuse$F.err; echo "resnet FUSE_BN=$F:"; python -c "import json;d=json.load(open('gpurun_out/${T}_bench_fuse$F.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
for F in 0 2; do
DEEPARDS_B200_FUSE_BN=$F timeout 600 python bench.py --no-cpu --backbone densenet18 > gpurun_out/${T}_bench_dense_fuse$F.json 2> gpurun_out/${T}_bench_dense_fuse$F.err; echo "densenet FUSE_BN=$F:"; python -c "import json;d=json.load(open('gpurun_out/${T}_bench_dense_fuse$F.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
