#!/bin/bash
mkdir -p gpurun_out
python tools/exp_fallback.py 2>&1 | grep "flow1080\|amp=0.60 \[" 
for m in exact nodes; do
  if [ $m = nodes ]; then export DVSG_TPS_NODES_FORCE=1; fi
  python bench.py --steps 10 --warmup 3 --workload cfg3 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$m cfg3 value %.1f frac %.3f kernel_ms %.4f mhz %s %s' % (d['value'], r['frac'], r['kernel_ms'], d['clocks']['sm_mhz'], r.get('kernel')))"
done
bash tools/gpu_prof2.sh bnod warp_bwd cfg3
