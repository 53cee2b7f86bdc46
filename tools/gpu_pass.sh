#!/bin/bash
# GPU pass: parity tests, smoke, microbench, bench, ncu launch list + full captures of the movement kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --maxfail=10 2>&1 | tail -60 > gpurun_out/pytest_e2e.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err
timeout 900 python bench.py --no-graphs --steps 60 --warmup 30 > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_microbench.csv python bench.py --microbench > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_move_kernel -s 20 -c 4 -o gpurun_out/prof_tma_move python bench.py --microbench > gpurun_out/ncu_tma.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 4 -c 2 -o gpurun_out/prof_gather_simt python bench.py --microbench > gpurun_out/ncu_simt.log 2>&1
fi
echo done
