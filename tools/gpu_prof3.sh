#!/bin/bash
# ncu --set full of the stem conv, the 8 ew_fused launches and the max-pool of one steady frame (graph nodes)
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 62 --skip-cpu-baseline --skip-e2e --skip-batched"
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:conv_stem -s 70 -c 1 -o gpurun_out/prof_stem_r01c -f $B > gpurun_out/ncu_stem.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:ew_fused -s 560 -c 8 -o gpurun_out/prof_ew_r01c -f $B > gpurun_out/ncu_ew.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:maxpool -s 70 -c 1 -o gpurun_out/prof_maxpool_r01c -f $B > gpurun_out/ncu_maxpool.log 2>&1
echo done
