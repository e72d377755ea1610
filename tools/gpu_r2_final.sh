#!/bin/bash
# round-2 evidence: ncu captures of the final kernels, launch lists, the bench lines of every workload, the kernel sweep
mkdir -p gpurun_out
bash tools/gpu_prof2.sh ${TAG:-r2p} warp_fwd cfg2 hd1080 mesh5 cfg5 cfg4
bash tools/gpu_prof2.sh ${TAG:-r2p} warp_bwd cfg3
unset DVSG_BENCH_MIN_S
: > gpurun_out/${TAG:-r2p}_bench_lines.jsonl
python bench.py --steps 20 --warmup 5 >> gpurun_out/${TAG:-r2p}_bench_lines.jsonl 2> gpurun_out/${TAG:-r2p}_bench.err
for wl in cfg1 cfg3 cfg3mask cfg3m5 cfg4 cfg5 hd1080 mesh5; do
  python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu >> gpurun_out/${TAG:-r2p}_bench_lines.jsonl 2>> gpurun_out/${TAG:-r2p}_bench.err
done
python tools/sweep.py all > gpurun_out/${TAG:-r2p}_sweep.txt 2>&1
tail -3 gpurun_out/${TAG:-r2p}_bench.err; wc -l gpurun_out/${TAG:-r2p}_bench_lines.jsonl; tail -5 gpurun_out/${TAG:-r2p}_sweep.txt
