#!/bin/bash
# host pipeline sweep: frames per chunk x staging slots, fp32 and uint8 host frames (cfg2: 64 x 720p)
export DVSG_BENCH_MIN_S=0.05
for c in 1 2 4 8; do for s in 2 3 4 6; do
  python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-chunk $c --e2e-slots $s --e2e-steps 8 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; u=d['e2e_u8']
print('chunk $c slots $s  e2e %.0f Mpix/s (%.1f GB/s over PCIe)  e2e_u8 %.0f Mpix/s (%.1f GB/s)' % (e['value'], e['value']*24e-3, u['value'], u['value']*6e-3))"
done; done
