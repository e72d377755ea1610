#!/bin/bash
mkdir -p gpurun_out
for PIPE in 0 1; do
DVSG_STRIP_PIPE=$PIPE timeout 1200 python -m pytest tests/test_gpu_forward.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_fwd_pipe$PIPE.log 2>&1
echo "pytest pipe=$PIPE exit $?"; tail -4 gpurun_out/pytest_fwd_pipe$PIPE.log
done
python tools/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"; tail -48 gpurun_out/sweep.log
CMD1="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD1 > gpurun_out/plain_cfg2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:warp_fwd -s 3 -c 1 -o gpurun_out/prof_tps_strip -f $CMD1 > gpurun_out/ncu_f_cfg2.log 2>&1
echo "ncu tps exit $?"
CMD2="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD2 > gpurun_out/plain_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:warp_fwd -s 3 -c 1 -o gpurun_out/prof_flow_strip -f $CMD2 > gpurun_out/ncu_f_cfg4.log 2>&1
echo "ncu flow exit $?"
