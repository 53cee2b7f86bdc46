#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/stage_probe.log
for cfg in "128 2" "128 3" "128 6" "64 2" "64 4"; do set -- $cfg
  echo "=== NTILE=$1 STAGES=$2" >> gpurun_out/stage_probe.log
  BC_CONV_NTILE=$1 BC_CONV_STAGES=$2 SPLIT=1 timeout 120 python tools/conv_bench.py 2>&1 | tail -1 >> gpurun_out/stage_probe.log
done
echo done
