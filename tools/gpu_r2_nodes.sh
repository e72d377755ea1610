#!/bin/bash
# node evaluation: parity gate + A/B against the per-pixel evaluation
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
: > gpurun_out/r2b_bench.jsonl
for wl in ${WLS:-cfg2 hd1080 mesh5 cfg5 cfg3 cfg1}; do
  for ex in 0 1; do
    if [ $ex = 1 ]; then export DVSG_TPS_EXACT=1; else unset DVSG_TPS_EXACT; fi
    echo "{\"tag\": \"$wl exact=$ex\"}" >> gpurun_out/r2b_bench.jsonl
    python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu --no-e2e --no-extras >> gpurun_out/r2b_bench.jsonl 2>> gpurun_out/r2b_bench.err
  done
done
unset DVSG_TPS_EXACT
grep -a "nodes-fp64\|pixels nodes\|passed\|failed\|FAILED" gpurun_out/r2b_pytest.log | cut -c1-260
python - <<'PY'
import json
tag=None
for l in open('gpurun_out/r2b_bench.jsonl'):
    d=json.loads(l)
    if 'tag' in d: tag=d['tag']; continue
    r=d['roofline']
    print(tag, 'value %.0f'%d['value'], 'frac %.3f'%r['frac'], 'kernel_ms %.4f'%r['kernel_ms'], 'fwd', r.get('forward_kernel',{}).get('kernel_ms'), d['clocks'])
PY
