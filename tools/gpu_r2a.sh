#!/bin/bash
# round-2 first GPU call: new fused conv+BN kernel tests, model parity, kbench, bench A/B
mkdir -p gpurun_out
T=${1:-r2a}
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -q --tb=short -p no:cacheprovider -x > gpurun_out/${T}_$name.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/${T}_$name.log | cut -c1-300; }
run convbn tests/test_kernels_gpu.py -m gpu -k "conv_bn"
timeout 600 python tools/kbench.py convbn > gpurun_out/${T}_kbench_convbn.txt 2>&1; cat gpurun_out/${T}_kbench_convbn.txt | tail -20
run model tests/test_model_parity_gpu.py tests/test_trainer_gpu.py -m gpu
DEEPARDS_B200_FUSE_BN=0 timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_nofuse.json 2> gpurun_out/${T}_bench_nofuse.err; head -c 400 gpurun_out/${T}_bench_nofuse.json; echo
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; head -c 400 gpurun_out/${T}_bench.json; echo; tail -3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --no-cpu --backbone densenet18 > gpurun_out/${T}_bench_dense.json 2> gpurun_out/${T}_bench_dense.err; head -c 400 gpurun_out/${T}_bench_dense.json; echo; tail -3 gpurun_out/${T}_bench_dense.err
