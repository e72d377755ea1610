#!/bin/bash
# last validation of the round: full GPU test-suite, smoke, the bench lines of every workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
: > gpurun_out/r2x_bench_lines.jsonl
python bench.py --gpus 1 --steps 20 --warmup 5 >> gpurun_out/r2x_bench_lines.jsonl 2> gpurun_out/r2x_bench.err; echo "default rc=$?"
for wl in cfg1 cfg3 cfg3mask cfg3m5 cfg4 cfg5 hd1080 mesh5; do
  python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu >> gpurun_out/r2x_bench_lines.jsonl 2>> gpurun_out/r2x_bench.err
done
tail -3 gpurun_out/r2x_bench.err; wc -l gpurun_out/r2x_bench_lines.jsonl
