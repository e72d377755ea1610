#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/quick_pytest.log 2>&1; tail -2 gpurun_out/quick_pytest.log
for wl in ${WLS:-hd1080 cfg2 mesh5 cfg5}; do
  python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$wl value %.0f frac %.3f kernel_ms %.4f mhz %s' % (d['value'], r['frac'], r['kernel_ms'], d['clocks']['sm_mhz']))"
done
