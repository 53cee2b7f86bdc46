#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q --maxfail=8 2>&1 | tail -30 > gpurun_out/pytest_conv.log
timeout 300 python -m pytest tests/test_gpu_e2e.py -m gpu -q --maxfail=8 2>&1 | tail -30 > gpurun_out/pytest_e2e.log
timeout 300 python tools/frame_breakdown.py > gpurun_out/frame_breakdown.log 2>&1
echo done
