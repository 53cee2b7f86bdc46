#!/bin/bash
mkdir -p gpurun_out
for b in 1 2 4 8; do timeout 600 python bench.py --skip-cpu-baseline --batch $b --steps 60 --warmup 30 > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; done
echo done
