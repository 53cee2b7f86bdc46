#!/bin/bash
# r02 profile set: launch list of steady frames (per graph node), ncu --set full of the share-dominant kernel
# (conv_igemm_persistent_kernel<128,5> on the layer3 shape, split-K) and of layer #20, CTA timelines.
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 62 --repeats 1 --skip-cpu-baseline --skip-e2e --skip-batched --skip-config4 --skip-rl --skip-microbench --skip-reference-gpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 3000 -c 1200 --csv --log-file gpurun_out/launches_bench_r02.csv $B > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 24 -c 3 -o gpurun_out/prof_conv_persist_r02 -f python tools/conv_bench.py > gpurun_out/ncu_conv_persist.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 8 -c 1 -o gpurun_out/prof_conv_l20_r02 -f python tools/conv_bench.py > gpurun_out/ncu_conv_l20.log 2>&1
python tools/cta_timeline.py > gpurun_out/cta_timeline_r02.txt 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
