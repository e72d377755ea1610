#!/bin/bash
# usage: gpu_prof2.sh <tag> <kernel-regex> <workload>...   -- plain run, launch list, one ncu --set full capture per workload
mkdir -p gpurun_out
TAG=$1; KREG=$2; shift 2
export DVSG_BENCH_MIN_S=0.02
for WL in "$@"; do
  CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras"
  $CMD > gpurun_out/${TAG}_plain_$WL.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain_$WL.log; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_$WL.csv $CMD > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:$KREG -s 4 -c 1 -o gpurun_out/${TAG}_prof_$WL -f $CMD > gpurun_out/${TAG}_ncu_$WL.log 2>&1
  echo "ncu exit $?"
  ncu -i gpurun_out/${TAG}_prof_$WL.ncu-rep --page raw --csv > gpurun_out/${TAG}_${WL}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_prof_$WL.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${WL}_sass.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_prof_$WL.ncu-rep
done
