#!/bin/bash
# usage: gpu_prof.sh <workload> <tag>   -- plain run, then one ncu --set full capture of the forward kernel
mkdir -p gpurun_out
WL=${1:-cfg2}; TAG=${2:-tile}
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_$WL.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$WL.log; exit 1; }
tail -c 600 gpurun_out/plain_$WL.log
ncu --set full --clock-control none --import-source on -k regex:warp_fwd -s 3 -c 1 -o gpurun_out/prof_${TAG}_$WL -f $CMD > gpurun_out/ncu_${TAG}_$WL.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_${TAG}_$WL.ncu-rep --page raw --csv > gpurun_out/${TAG}_${WL}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}_$WL.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${WL}_sass.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}_$WL.ncu-rep
