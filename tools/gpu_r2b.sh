#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2b}
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -q --tb=short -p no:cacheprovider > gpurun_out/${T}_$name.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/${T}_$name.log | cut -c1-300; }
run convbn tests/test_kernels_gpu.py -m gpu -k "conv_bn or tcgen05" --maxfail=5
timeout 600 python tools/kbench.py convbn > gpurun_out/${T}_kbench_convbn.txt 2>&1; tail -20 gpurun_out/${T}_kbench_convbn.txt
echo "--- conv, shared loop off"; KBENCH_SHARED=0 timeout 600 python tools/kbench.py conv > gpurun_out/${T}_kbench_conv_shared0.txt 2>&1; grep -v wgrad gpurun_out/${T}_kbench_conv_shared0.txt
echo "--- conv, shared loop default (C>128)"; timeout 600 python tools/kbench.py conv > gpurun_out/${T}_kbench_conv.txt 2>&1; grep -v wgrad gpurun_out/${T}_kbench_conv.txt

run model tests/test_model_parity_gpu.py tests/test_trainer_gpu.py -m gpu
for F in 0 2 1; do
DEEPARDS_B200_FUSE_BN=$F timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_fuse$F.json 2> gpurun_out/${T}_bench_fuse$F.err; echo "FUSE_BN=$F"; head -c 330 gpurun_out/${T}_bench_fuse$F.json | tail -c 130; echo; tail -2 gpurun_out/${T}_bench_fuse$F.err
done
DEEPARDS_B200_TC_DEBUG="10=0" DEEPARDS_B200_FUSE_BN=0 timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_old.json 2> gpurun_out/${T}_bench_old.err; echo "all old"; head -c 330 gpurun_out/${T}_bench_old.json | tail -c 130; echo
for F in 0 2 1; do
DEEPARDS_B200_FUSE_BN=$F timeout 600 python bench.py --no-cpu --backbone densenet18 > gpurun_out/${T}_bench_dense_fuse$F.json 2> gpurun_out/${T}_bench_dense_fuse$F.err; echo "dense FUSE_BN=$F"; head -c 330 gpurun_out/${T}_bench_dense_fuse$F.json | tail -c 130; echo; tail -2 gpurun_out/${T}_bench_dense_fuse$F.err
done
