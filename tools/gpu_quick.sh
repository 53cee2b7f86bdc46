#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/tma_sweep.py > gpurun_out/tma_sweep.log 2>&1
echo done
