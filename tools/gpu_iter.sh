#!/bin/bash
# one iteration on the GPU box: forward (+ optionally backward) parity tests, then the CUDA-event sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_forward.py ${TESTS_EXTRA} -m gpu -q -x --timeout 300 > gpurun_out/pytest_fwd.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_fwd.log
timeout 900 python tools/sweep.py ${1:-all} > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"; cat gpurun_out/sweep.log
