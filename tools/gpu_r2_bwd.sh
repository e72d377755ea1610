#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_backward.py tests/test_gpu_losses.py tests/test_gpu_forward.py -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1
tail -4 gpurun_out/r2e_pytest.log
: > gpurun_out/r2e_bench.jsonl
for lib in default small; do
  if [ $lib = small ]; then export DVSG_LIB=$PWD/build_ab/libdvsg_small.so; else unset DVSG_LIB; fi
  for wl in cfg3 cfg3mask; do
    echo "{\"tag\": \"$wl lib=$lib\"}" >> gpurun_out/r2e_bench.jsonl
    python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu --no-e2e --no-extras >> gpurun_out/r2e_bench.jsonl 2>> gpurun_out/r2e_bench.err
  done
done
unset DVSG_LIB
python - <<'PY'
import json
tag=None
for l in open('gpurun_out/r2e_bench.jsonl'):
    d=json.loads(l)
    if 'tag' in d: tag=d['tag']; continue
    r=d['roofline']
    print(tag, 'value %.0f'%d['value'], 'ms/pass %.4f'%d['timing']['ms_per_pass'], 'bwd frac %.3f'%r['frac'], 'bwd_ms %.4f'%r['kernel_ms'], 'fwd_ms %.4f'%r['forward_kernel']['kernel_ms'], 'fwd frac %.3f'%r['forward_kernel']['frac'])
PY
tail -3 gpurun_out/r2e_bench.err
