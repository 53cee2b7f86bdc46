#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_policy.py tests/test_gpu_e2e.py -m gpu -q --maxfail=8 2>&1 | tail -30 > gpurun_out/pytest_policy.log
timeout 600 python bench.py --skip-cpu-baseline --policy rl_semseg > gpurun_out/bench_rl.json 2> gpurun_out/bench_rl.err
echo done
