#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 3000 -c 1000 --csv --log-file gpurun_out/launches_bench_r01b.csv python bench.py --steps 40 --warmup 62 --skip-cpu-baseline --skip-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 2 -o gpurun_out/prof_conv_l20 python tools/conv_bench.py > gpurun_out/ncu_conv_l20.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_stem -c 1 -o gpurun_out/prof_conv_stem python -m pytest tests/test_gpu_conv.py -m gpu -q -k "stem and 128-64" > gpurun_out/ncu_stem.log 2>&1
echo done
