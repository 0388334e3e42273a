#!/bin/bash
# compute-sanitizer over the per-kernel GPU tests (race / memory checking).
#   tools/sanitize.sh memcheck|racecheck|synccheck|initcheck [pytest -k expression]
# ONE tool per gpurun call (B200_PROFILING.md: several tools in one call have left GPUs unusable), on the small kernel
# cases only.  The run is bounded by `timeout`; output goes to gpurun_out/sanitize_<tool>.log.
TOOL=${1:-memcheck}
EXPR=${2:-"tcgen05 or conv_bn or gbn or stem"}
mkdir -p gpurun_out
case "$TOOL" in memcheck|racecheck|synccheck|initcheck) ;; *) echo "unknown tool $TOOL"; exit 2;; esac
# the plain run first: a faulting program must not be put under the sanitizer
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "$EXPR" > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool "$TOOL" --error-exitcode 3 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "$EXPR" > gpurun_out/sanitize_${TOOL}.log 2>&1
RC=$?
echo "compute-sanitizer --tool $TOOL exit $RC"
grep -E "ERROR SUMMARY|passed|failed|Invalid|hazard" gpurun_out/sanitize_${TOOL}.log | tail -8
exit $RC
