#!/bin/bash
# usage: gpu_ab_modes.sh <workload>...  -- bench each workload with the tile-node evaluation and with DVSG_TPS_EXACT=1
mkdir -p gpurun_out
: > gpurun_out/ab_modes.jsonl
for wl in "$@"; do
  for ex in 0 1; do
    if [ $ex = 1 ]; then export DVSG_TPS_EXACT=1; else unset DVSG_TPS_EXACT; fi
    echo "{\"tag\": \"$wl exact=$ex\"}" >> gpurun_out/ab_modes.jsonl
    python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu --no-e2e --no-extras >> gpurun_out/ab_modes.jsonl 2>> gpurun_out/ab_modes.err
  done
done
unset DVSG_TPS_EXACT
python - <<'PY'
import json
tag=None
for l in open('gpurun_out/ab_modes.jsonl'):
    d=json.loads(l)
    if 'tag' in d: tag=d['tag']; continue
    r=d['roofline']; f=r.get('forward_kernel')
    print(tag, 'value %.0f'%d['value'], 'ms/pass %.4f'%d['timing']['ms_per_pass'], 'frac %.3f'%r['frac'], 'kernel_ms %.4f'%r['kernel_ms'], ('fwd_ms %.4f frac %.3f'%(f['kernel_ms'], f['frac'])) if f else '', d['clocks']['sm_mhz'])
PY
