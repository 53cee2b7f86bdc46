#!/bin/bash
# r01c profile set: launch list of a steady frame + ncu --set full of the roofline conv (#20), the persistent conv
# (layer2 shape) and the stem
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 62 --skip-cpu-baseline --skip-e2e --skip-batched"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 3000 -c 1200 --csv --log-file gpurun_out/launches_bench_r01c.csv python bench.py --steps 40 --warmup 62 --skip-cpu-baseline --skip-e2e --skip-batched > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 8 -c 1 -o gpurun_out/prof_conv_l20_r01c -f python tools/conv_bench.py > gpurun_out/ncu_conv_l20.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 8 -c 1 -o gpurun_out/prof_conv_persist_r01c -f python tools/conv_bench.py > gpurun_out/ncu_conv_persist.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:conv_stem -s 70 -c 1 -o gpurun_out/prof_stem_r01c2 -f $B > gpurun_out/ncu_stem.log 2>&1
echo done
