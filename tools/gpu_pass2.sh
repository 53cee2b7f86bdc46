#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --maxfail=10 2>&1 | tail -30 > gpurun_out/pytest_kernels.log
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q --maxfail=6 2>&1 | tail -60 > gpurun_out/pytest_conv.log
BLOCKCOPY_FUSED_CONV=0 timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/pytest_e2e_nofuse.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/pytest_e2e.log
timeout 600 python bench.py --microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err
BLOCKCOPY_FUSED_CONV=0 timeout 600 python bench.py --no-graphs --steps 60 --warmup 30 --skip-cpu-baseline > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err
timeout 600 python bench.py --steps 90 --warmup 60 --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo done
