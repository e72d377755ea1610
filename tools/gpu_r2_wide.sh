#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_forward.py tests/test_gpu_properties.py tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; tail -4 gpurun_out/r2l_pytest.log
echo "--- staged"; python tools/sweep.py c18
echo "--- direct (DVSG_WIDE_DIRECT=1)"; DVSG_WIDE_DIRECT=1 python tools/sweep.py c18
