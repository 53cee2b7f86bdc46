#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --maxfail=8 2>&1 | tail -40 > gpurun_out/pytest_conv.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --maxfail=8 2>&1 | tail -8 > gpurun_out/pytest_kernels.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --maxfail=8 2>&1 | tail -40 > gpurun_out/pytest_e2e.log
BLOCKCOPY_LAZY=0 timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --maxfail=8 2>&1 | tail -10 > gpurun_out/pytest_e2e_nolazy.log
timeout 600 python tools/frame_breakdown.py > gpurun_out/frame_breakdown.log 2>&1
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo done
