#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke, the reference arm, the own arm (default flags)
mkdir -p gpurun_out
T0=$(date +%s)
python -m pytest tests/ -x -q -m gpu > gpurun_out/end_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/end_pytest.log
T0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; echo "smoke $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
python bench.py --impl reference > gpurun_out/end_bench_ref.json 2> gpurun_out/end_bench_ref.err; echo "reference arm rc=$? $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
python bench.py > gpurun_out/end_bench.json 2> gpurun_out/end_bench.err; echo "own arm rc=$? $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
for f in ('gpurun_out/end_bench_ref.json', 'gpurun_out/end_bench.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d.get('metric'), round(d['value'], 1), d['unit'], 'e2e', round(d['e2e']['value'], 1), 'frac', d.get('roofline', {}).get('frac'),
          'north_star', (d.get('north_star') or {}).get('roofline', {}).get('frac') if isinstance(d.get('north_star'), dict) else None, d.get('clocks'))
PY
