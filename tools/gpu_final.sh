#!/bin/bash
# Round evidence in one call: tests + smoke + bench lines (gpu_check.sh), sweeps, then launch lists and ncu captures (gpu_profile_all.sh)
bash tools/gpu_check.sh
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench exit $?"; tail -c 600 gpurun_out/bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; tail -c 400 gpurun_out/bench_reference.json
python tools/sweep.py all > gpurun_out/sweep_final.log 2>&1; echo "sweep exit $?"
(python tools/sweep.py bwd; python tools/exp_flow_bwd.py; python tools/sweep.py c18; python tools/exp_bwd_split.py) > gpurun_out/sweep_bwd.log 2>&1; echo "sweep bwd exit $?"
bash tools/gpu_profile_all.sh
