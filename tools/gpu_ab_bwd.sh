#!/bin/bash
# A/B of backward-kernel builds (build_ab/lib_<tag>.so): parity tests, then the cfg3 bench line and the backward sweep
for tag in "$@"; do
  echo "=== $tag"
  export DVSG_LIB=$PWD/build_ab/lib_$tag.so
  python -m pytest tests/test_gpu_backward.py tests/test_gpu_losses.py -m gpu -q -x --timeout 300 2>&1 | tail -2
  python bench.py --workload cfg3 --steps 100 --warmup 5 --no-cpu 2>&1 | grep -o "\"kernel_ms\": [0-9.]*\|\"ms_per_step\": [0-9.]*" | tr "\n" " "; echo
  python tools/sweep.py bwd 2>&1 | cut -c1-110
done
