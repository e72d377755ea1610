#!/bin/bash
for c in 1 2 4 8; do for s in 3 4 6 8; do
  python bench.py --steps 5 --warmup 3 --no-cpu --e2e-chunk $c --e2e-slots $s --e2e-steps 8 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('chunk $c slots $s  e2e %.0f  e2e_u8 %.0f Mpix/s' % (d['e2e']['value'], d['e2e_u8']['value']))"
done; done
