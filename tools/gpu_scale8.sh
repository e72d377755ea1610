#!/bin/bash
# 8-GPU evidence: the driver's own launch line (torchrun, one rank per GPU) for the default workload, then cfg5 and cfg4
mkdir -p gpurun_out
N=${1:-8}
for WL in cfg2 cfg5 cfg4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --workload $WL --no-cpu > gpurun_out/scale${N}_$WL.json 2> gpurun_out/scale${N}_$WL.err
  echo "$WL exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"e2e": {[^}]*}' gpurun_out/scale${N}_$WL.json | head -4
done
