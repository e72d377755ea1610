#!/bin/bash
# bench lines of every workload (no ncu): TAG names the output
mkdir -p gpurun_out
unset DVSG_BENCH_MIN_S
: > gpurun_out/${TAG:-fin}_bench_lines.jsonl
python bench.py --steps 20 --warmup 5 >> gpurun_out/${TAG:-fin}_bench_lines.jsonl 2> gpurun_out/${TAG:-fin}_bench.err
for wl in cfg1 cfg3 cfg3mask cfg3m5 cfg4 cfg5 hd1080 mesh5; do
  python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu >> gpurun_out/${TAG:-fin}_bench_lines.jsonl 2>> gpurun_out/${TAG:-fin}_bench.err
done
wc -l gpurun_out/${TAG:-fin}_bench_lines.jsonl; tail -2 gpurun_out/${TAG:-fin}_bench.err
