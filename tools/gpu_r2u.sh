#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2u}
timeout 1500 python -m pytest tests/ -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('module_path',d.get('module_path')); print('densenet', d['densenet18']['value'])
PY
