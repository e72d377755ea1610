#!/bin/bash
export DVSG_BENCH_MIN_S=0.02
for m in one two; do
  if [ $m = one ]; then export DVSG_TPS_ONE_LEVEL=1; else unset DVSG_TPS_ONE_LEVEL; fi
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,launch__grid_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:warp_fwd_tile -s 3 -c 1 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras 2>&1 | grep -E "inst_executed|time_duration|grid_size|occupancy_limit|issue_active" | sed "s/^/$m: /"
done
