#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2u_pytest.log 2>&1
tail -3 gpurun_out/r2u_pytest.log; grep -a "node-mode grad_T\|FAILED" gpurun_out/r2u_pytest.log | head
python -c "import __graft_entry__ as g; g.smoke()"
python tools/sweep.py c18 > gpurun_out/r2u_sweep_c18.txt 2>&1; cat gpurun_out/r2u_sweep_c18.txt
