#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/split_sweep.log
for mc in 48 96 160; do for tg in 240 320 480; do
  echo "MAX_CTAS=$mc TARGET=$tg" >> gpurun_out/split_sweep.log
  BC_SPLIT_MAX_CTAS=$mc BC_SPLIT_TARGET=$tg SPLIT=1 timeout 120 python tools/conv_bench.py 2>&1 | tail -1 >> gpurun_out/split_sweep.log
done; done
echo done
