#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "tcgen05" 2>&1 | tail -4
timeout 300 python bench.py --no-extra --no-cpu --steps 60 --warmup 10 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; python -c "
import json; d=json.load(open('gpurun_out/p_bench.json')); print(d['value'], d['ms_per_step'])"
