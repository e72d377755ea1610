#!/bin/bash
# ncu --set full of the forward kernel at cfg5 (4K, 16x16 mesh) and mesh5 (720p, 5x5 mesh), after plain runs of the same commands
mkdir -p gpurun_out
for WL in cfg5 mesh5; do
  CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --no-e2e"
  $CMD > gpurun_out/plain_$WL.log 2>&1 || { echo "plain $WL failed"; tail -3 gpurun_out/plain_$WL.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:warp_fwd -s 3 -c 1 -o gpurun_out/prof_final_$WL -f $CMD > gpurun_out/ncu_final_$WL.log 2>&1
  echo "ncu full $WL exit $?"
  ncu -i gpurun_out/prof_final_$WL.ncu-rep --page raw --csv > gpurun_out/final_${WL}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_final_$WL.ncu-rep --page source --csv --print-source sass > gpurun_out/final_${WL}_sass.csv 2>/dev/null
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$WL.csv $CMD > /dev/null 2>&1
  echo "launch list $WL exit $?"
done
