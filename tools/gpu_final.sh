#!/bin/bash
# Round-end validation on one B200: GPU test-suite, smoke(), the default bench line, the reference (CPU) arm.
mkdir -p gpurun_out
T=${1:-final}
timeout 1700 python -m pytest tests/ -m gpu -q -p no:cacheprovider 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "reference arm exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',d['e2e']['value'],'sustained',d['sustained']['value']); print('module_path',d.get('module_path')); print('densenet', d['densenet18']['value']); print('cpu', d['cpu_baseline']); print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('per_kernel','breakdown_ms_per_step')})
r=json.load(open('gpurun_out/${T}_bench_ref.json')); print('ref', {k:r[k] for k in ('impl','value','unit','ms_per_step') if k in r})
PY
