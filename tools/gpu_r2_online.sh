#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1
tail -4 gpurun_out/r2h_pytest.log
python bench.py --steps 20 --warmup 5 --workload cfg1 --no-cpu > gpurun_out/r2h_cfg1.json 2> gpurun_out/r2h_cfg1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_cfg1.json').read().strip().splitlines()[-1])
print('cfg1 value %.0f Mpix/s'%d['value'], 'us per frame %.2f'%(d['timing']['ms_per_pass']*1e3), 'kernel_us %.2f'%(d['roofline']['kernel_ms']*1e3), 'launches', d['gpu_launches'], 'passes', d['timing']['passes_per_step'], 'e2e', d['e2e'] and d['e2e']['value'])
PY
tail -2 gpurun_out/r2h_cfg1.err
