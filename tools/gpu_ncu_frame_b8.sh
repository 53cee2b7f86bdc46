#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 3000 -c 1200 --csv --log-file gpurun_out/launches_bench_b8.csv python bench.py --batch 8 --steps 40 --warmup 62 --skip-cpu-baseline --skip-e2e --skip-batched > gpurun_out/ncu_bench_b8.log 2>&1
echo done
