#!/bin/bash
# ncu evidence for the dominant kernels (one plain run first, then the profiled run).
mkdir -p gpurun_out
python tools/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"; cat gpurun_out/sweep.log
CMD1="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu --no-e2e"
CMD2="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD1 > gpurun_out/plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2.csv $CMD1 > gpurun_out/ncu_l_cfg2.log 2>&1
echo "launch list exit $?"
$CMD1 > gpurun_out/plain_cfg2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:warp_fwd -s 3 -c 2 -o gpurun_out/prof_tps_cfg2 -f $CMD1 > gpurun_out/ncu_f_cfg2.log 2>&1
echo "ncu full tps exit $?"
$CMD2 > gpurun_out/plain_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:warp_fwd -s 3 -c 2 -o gpurun_out/prof_flow_cfg4 -f $CMD2 > gpurun_out/ncu_f_cfg4.log 2>&1
echo "ncu full flow exit $?"
ls -la gpurun_out/*.ncu-rep
