#!/bin/bash
# scaling evidence at N GPUs with the driver's own launch line (default workload only)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu > gpurun_out/scale${N}_cfg2.json 2> gpurun_out/scale${N}_cfg2.err
echo "exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/scale${N}_cfg2.json | head -3
