#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_policy.py tests/test_gpu_e2e.py -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_rl.log
timeout 600 python bench.py --policy rl_semseg --skip-cpu-baseline --skip-e2e --skip-batched > gpurun_out/bench_rl.json 2> gpurun_out/bench_rl.err
echo done
