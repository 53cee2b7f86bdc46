#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --policy rl_semseg --shared-policy --steps 90 --warmup 30 --skip-cpu-baseline --skip-e2e --skip-batched > gpurun_out/bench_2gpu_rl_shared.json 2> gpurun_out/bench_2gpu_rl_shared.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --policy rl_semseg --steps 90 --warmup 30 --skip-cpu-baseline --skip-e2e --skip-batched > gpurun_out/bench_2gpu_rl.json 2> gpurun_out/bench_2gpu_rl.err
echo done
