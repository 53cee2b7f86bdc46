#!/bin/bash
# BC_CONV_DEBUG bits: 4 no MMAs, 8 no A loads, 16 no B loads (after the prologue), 32 producers ignore the empty
# barriers, 64 software arrive instead of tcgen05.commit
mkdir -p gpurun_out
: > gpurun_out/stage_probe.log
for dbg in 28 60 92 124 32 40 48; do
  echo "=== DEBUG=$dbg (NTILE=128 STAGES=3)" >> gpurun_out/stage_probe.log
  BC_CONV_DEBUG=$dbg BC_CONV_NTILE=128 BC_CONV_STAGES=3 timeout 120 python tools/cta_timeline.py 2>&1 | grep -A2 "layer2" | grep -v "^--" >> gpurun_out/stage_probe.log
done
echo done
