#!/bin/bash
# r02b profile set (final build of round 2): launch list of steady frames (per graph node) and ncu --set full of the
# share-dominant kernel (conv_igemm_persistent_kernel<128,5>, TMA-store epilogue) on the layer2 / layer3 / layer4 shapes.
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 62 --repeats 1 --skip-cpu-baseline --skip-e2e --skip-batched --skip-config4 --skip-rl --skip-microbench --skip-reference-gpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 3000 -c 1200 --csv --log-file gpurun_out/launches_bench_r02b.csv $B > gpurun_out/ncu_bench_r02b.log 2>&1
tail -2 gpurun_out/ncu_bench_r02b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 24 -c 9 -o gpurun_out/prof_conv_persist_r02b -f python tools/conv_bench.py > gpurun_out/ncu_conv_persist_r02b.log 2>&1
tail -2 gpurun_out/ncu_conv_persist_r02b.log
python tools/cta_timeline.py > gpurun_out/cta_timeline_r02b.txt 2>&1
ls -la gpurun_out/*r02b*
echo done
