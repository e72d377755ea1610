#!/bin/bash
# First-contact GPU run: tests (full log), smoke, short bench for each workload.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=12 --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
for wl in cfg2 hd1080 mesh5 cfg4 cfg5 cfg3 cfg1; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "bench $wl exit $?"; tail -c 1500 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
