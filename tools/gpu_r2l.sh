#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2l}
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
echo "--- conv kbench, old kernels (shared loop off)"; KBENCH_SHARED=0 timeout 300 python tools/kbench.py conv 2>&1 | tee gpurun_out/${T}_kbench_conv_old.txt
echo "--- conv kbench, shared loop wide tile for C>128"; KBENCH_CB_WIDE=1 timeout 300 python tools/kbench.py conv 2>&1 | grep -v wgrad | tail -4
timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_trainer_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
DEEPARDS_B200_TC_DEBUG="10=0" DEEPARDS_B200_FUSE_BN=0 timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_oldpath.json 2> gpurun_out/${T}_bench_oldpath.err; echo "old path:"; python -c "import json;d=json.load(open('gpurun_out/${T}_bench_oldpath.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'])"
DEEPARDS_B200_TC_DEBUG="13=1" DEEPARDS_B200_FUSE_BN=0 timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_wide.json 2> gpurun_out/${T}_bench_wide.err; echo "shared wide tile, no fused bn:"; python -c "import json;d=json.load(open('gpurun_out/${T}_bench_wide.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'])"
DEEPARDS_B200_TC_DEBUG="13=1" DEEPARDS_B200_FUSE_BN=2 timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_fuse2.json 2> gpurun_out/${T}_bench_fuse2.err; echo "shared wide tile + fused bn (mode 2):"; python -c "import json;d=json.load(open('gpurun_out/${T}_bench_fuse2.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'])"
DEEPARDS_B200_TC_DEBUG="10=0" DEEPARDS_B200_FUSE_BN=0 timeout 600 python bench.py --no-cpu --backbone densenet18 > gpurun_out/${T}_bench_dense_old.json 2> gpurun_out/${T}_bench_dense_old.err; echo "densenet old path:"; python -c "import json;d=json.load(open('gpurun_out/${T}_bench_dense_old.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'])"
DEEPARDS_B200_FUSE_BN=2 timeout 600 python bench.py --no-cpu --backbone densenet18 > gpurun_out/${T}_bench_dense_fuse2.json 2> gpurun_out/${T}_bench_dense_fuse2.err; echo "densenet fused bn:"; python -c "import json;d=json.load(open('gpurun_out/${T}_bench_dense_fuse2.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'])"
