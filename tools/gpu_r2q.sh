#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2q}
timeout 1200 python -m pytest tests/test_model_parity_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/${T}_parity.log 2>&1; echo "exit $?"; grep -E "bf16|worst grad|config1|passed|failed|Error" gpurun_out/${T}_parity.log | cut -c1-250 | head -30
grep -E "worst grad|config1|bf16 B=256|trajectory|bf16:" gpurun_out/${T}_parity.log > gpurun_out/${T}_parity_report.txt
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e',d['e2e']); print('sustained',d.get('sustained')); print('module_path',d.get('module_path'))
dn=d.get('densenet18'); print('densenet18', {k:dn[k] for k in ('value','ms_per_step','gpu_launches')} if dn else None, dn and dn.get('roofline',{}).get('frac'))
r=d['roofline']; print('roofline',{k:r[k] for k in ('kernel','bound','achieved','peak','frac','share_of_step')}); print(r['breakdown_ms_per_step'])
print('cpu', d.get('cpu_baseline'))
PY
