#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py -m gpu -q -x -s 2>&1 | grep -i "cfg5\|16x16\|m16\|passed\|failed\|error" | tail -12
for m in one two; do
  if [ $m = one ]; then export DVSG_TPS_ONE_LEVEL=1; else unset DVSG_TPS_ONE_LEVEL; fi
  python bench.py --steps 10 --warmup 3 --workload cfg5 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$m-level cfg5 value %.0f frac %.3f kernel_ms %.4f mhz %s' % (d['value'], r['frac'], r['kernel_ms'], d['clocks']['sm_mhz']))"
done
