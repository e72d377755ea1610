#!/bin/bash
# the driver's own launch lines at N GPUs, both arms, default flags
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/final_n${N}_ref.json 2> gpurun_out/final_n${N}_ref.err
echo "reference arm exit $? after $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_n${N}.json 2> gpurun_out/final_n${N}.err
echo "own arm exit $? after $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/final_n${N}.err
python - <<PY
import json
for f in ('gpurun_out/final_n${N}_ref.json', 'gpurun_out/final_n${N}.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('unit'), 'e2e', d.get('e2e', {}).get('value'), 'n', d.get('n_gpus'), 'frac', d.get('roofline', {}).get('frac'), 'clocks', d.get('clocks'))
        for k in ('strong',):
            if k in d: print('  strong', json.dumps(d[k])[:600])
    except Exception as e:
        print(f, 'unreadable', e)
PY
