#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_forward.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/fb_pytest.log 2>&1; tail -3 gpurun_out/fb_pytest.log
python tools/exp_fallback.py > gpurun_out/fb_ab.txt 2>&1; cat gpurun_out/fb_ab.txt
