#!/bin/bash
# round 2, first GPU pass: full GPU test-suite, racecheck of the shared-mesh solve, the new bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -80 > gpurun_out/r2a_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python - > gpurun_out/r2a_racecheck.log 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
from coupe.dvsg_b200 import _lib
from oracle import dvsg_oracle as O
lib = _lib.load()
for n in (5, 4):
    pn, B = n * n, 40
    mesh = torch.from_numpy(O.regular_mesh(n, n)).cuda()
    tgt = (mesh[None] + (torch.rand(B, pn, 2, device='cuda') - 0.5) * 0.2).contiguous()
    T = torch.empty(B, 2, pn + 3, device='cuda')
    assert lib.dvsg_tps_solve(mesh.data_ptr(), 0, tgt.data_ptr(), T.data_ptr(), B, pn, 0, 0, 0) == 0
    g = torch.empty(B, pn, 2, device='cuda')
    assert lib.dvsg_tps_solve_bwd(mesh.data_ptr(), 0, T.data_ptr(), g.data_ptr(), B, pn, 0, 0, 0) == 0
torch.cuda.synchronize()
print('racecheck run done')
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
python bench.py --steps 20 --warmup 5 --workload cfg3 --no-cpu > gpurun_out/r2a_bench_cfg3.json 2>> gpurun_out/r2a_bench.err
python bench.py --steps 20 --warmup 5 --workload mesh5 --no-cpu --no-e2e > gpurun_out/r2a_bench_mesh5.json 2>> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; tail -3 gpurun_out/r2a_racecheck.log; cut -c1-600 gpurun_out/r2a_bench.json
