#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/frame_breakdown.py > gpurun_out/frame_breakdown.log 2>&1
echo done
