#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 62 --skip-cpu-baseline --skip-e2e --skip-batched"
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:head_1x1 -s 70 -c 1 -o gpurun_out/prof_head -f $B > gpurun_out/ncu_head.log 2>&1
echo done
