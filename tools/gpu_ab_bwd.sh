#!/bin/bash
for tag in "$@"; do
  echo "=== $tag"
  DVSG_LIB=$PWD/build_ab/lib_$tag.so python bench.py --workload cfg3 --steps 100 --warmup 5 --no-cpu 2>&1 | grep -o "\"kernel_ms\": [0-9.]*\|\"ms_per_step\": [0-9.]*" | tr "\n" " "; echo
done
