#!/bin/bash
# node mode vs exact at the 288x512 training shape (DVSG_TPS_NODES_FORCE lifts the pixel floor of tps_nodes_ok)
for wl in cfg3 cfg3m5; do
for m in exact nodes; do
  if [ $m = nodes ]; then export DVSG_TPS_NODES_FORCE=1; else unset DVSG_TPS_NODES_FORCE; fi
  python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$m $wl value %.1f frac %.3f bwd_ms %.4f fwd_ms %.4f mhz %s' % (d['value'], r['frac'], r['kernel_ms'], r['forward_kernel']['kernel_ms'], d['clocks']['sm_mhz']))"
done; done
