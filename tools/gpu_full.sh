#!/bin/bash
# what the driver runs at round end, plus the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python tools/frame_breakdown.py > gpurun_out/frame_breakdown.log 2>&1
echo done
