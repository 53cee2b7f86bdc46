#!/bin/bash
# r01d: ncu --set full of bc_bn_stats (level-0 launch of the fused policy trunk: 1x64x256x512 fp16) and bc_pack_params
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_stats -s 12 -c 1 -o gpurun_out/prof_bn_stats_r01d -f python tools/policy_bench.py > gpurun_out/ncu_bn_stats.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pack_params -s 1 -c 1 -o gpurun_out/prof_pack_r01d -f python tools/policy_bench.py > gpurun_out/ncu_pack.log 2>&1
echo done
