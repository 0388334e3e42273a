#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_extra_gpu.py -q -p no:cacheprovider -x -s -k "flat_batch" 2>&1 | grep "flat batch\|passed\|failed\|Error" | head -20
timeout 120 python tools/kbench.py stem 2>&1 | tail -2
