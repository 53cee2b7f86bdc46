#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q --maxfail=5 2>&1 | tail -12 > gpurun_out/pytest_conv.log
(for d in 0 2; do SPLIT=0 BC_CONV_DEBUG=$d python tools/conv_bench.py; done) > gpurun_out/conv_bench.log 2>&1
timeout 300 python tools/frame_breakdown.py > gpurun_out/frame_breakdown.log 2>&1
echo done
