#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_cfg3.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_cfg3.log; exit 1; }
tail -c 900 gpurun_out/plain_cfg3.log
ncu --set full --clock-control none --import-source on -k regex:warp_bwd -s 3 -c 1 -o gpurun_out/prof_bwd_cfg3 -f $CMD > gpurun_out/ncu_bwd_cfg3.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_bwd_cfg3.ncu-rep --page raw --csv > gpurun_out/bwd_cfg3_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_bwd_cfg3.ncu-rep --page source --csv --print-source sass > gpurun_out/bwd_cfg3_sass.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg3.csv $CMD > /dev/null 2>&1
echo "launch list exit $?"
