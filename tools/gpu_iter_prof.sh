#!/bin/bash
# tests + sweep, then one ncu --set full capture of the forward kernel at <workload> (default cfg2)
bash tools/gpu_iter.sh ${1:-tps}
bash tools/gpu_prof.sh ${2:-cfg2} ${3:-iter} > gpurun_out/prof_iter.log 2>&1; tail -3 gpurun_out/prof_iter.log
