#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py tests/test_gpu_spp.py -m gpu -q --maxfail=5 2>&1 | tail -8 > gpurun_out/pytest_conv.log
(SPLIT=1 timeout 120 python tools/conv_bench.py) > gpurun_out/conv_bench.log 2>&1
timeout 300 python tools/frame_breakdown.py > gpurun_out/frame_breakdown.log 2>&1
echo done
