#!/bin/bash
# per-kernel launch list of the graphed frame loop + full captures of the dominant kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --maxfail=10 2>&1 | tail -5 > gpurun_out/pytest_kernels.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 4000 -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 40 --warmup 62 --skip-cpu-baseline --skip-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:conv_igemm -s 60 -c 6 -o gpurun_out/prof_conv python bench.py --steps 10 --warmup 35 --skip-cpu-baseline --skip-e2e > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_move_kernel -s 20 -c 4 -o gpurun_out/prof_tma_move python bench.py --microbench > gpurun_out/ncu_tma.log 2>&1
echo done
