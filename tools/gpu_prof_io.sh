#!/bin/bash
# r01d: ncu --set full of bc_frame_from_u8 / bc_upsample_argmax at 1024x2048 (third launch of each)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_from_u8 -s 2 -c 1 -o gpurun_out/prof_io_u8_r01d -f python tools/io_kernels.py > gpurun_out/ncu_io_u8.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:upsample_argmax -s 2 -c 1 -o gpurun_out/prof_io_argmax_r01d -f python tools/io_kernels.py > gpurun_out/ncu_io_argmax.log 2>&1
echo done
