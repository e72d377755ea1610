#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_forward.py -m gpu -q -k "host_pipeline or online" > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
export DVSG_BENCH_MIN_S=0.05
for wl in cfg2 cfg4 cfg5 cfg1; do
 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --workload $wl --e2e-steps 8 2>>gpurun_out/r2j.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; u=d.get('e2e_u8')
print('$wl e2e %.0f Mpix/s streaming (%.1f GB/s), blocking %.0f' % (e['value'], e['pcie_gbs'], e['blocking_value']), ('| u8 %.0f (%.1f GB/s) blocking %.0f' % (u['value'], u['pcie_gbs'], u['blocking_value'])) if u else '', '| us/pass %.2f' % (d['timing']['ms_per_pass']*1e3))"
done
tail -3 gpurun_out/r2j.err
