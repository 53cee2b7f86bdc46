#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_e2e.log
echo done
