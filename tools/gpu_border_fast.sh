#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_properties.py tests/test_gpu_losses.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/sweep.py bwd 2>&1 | grep "tf_warp\|tps bwd"
