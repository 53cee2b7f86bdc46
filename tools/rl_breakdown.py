#!/usr/bin/env python
"""Where does an rl_semseg frame go?  Wall time per phase with a device sync after each (so phases do not overlap;
the sum is larger than the pipelined frame time)."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))
import torch
import blockcopy
from blockcopy.utils.profiler import timings
from consumers.clips import synthetic_clip
from consumers.swiftnet_rn18 import build_swiftnet_rn18
import bench

random.seed(0); torch.manual_seed(0)
settings = bench.default_block_settings() if hasattr(bench, "default_block_settings") else None
if settings is None:
    from blockcopy.core.argparser import default_settings
    settings = default_settings()
settings.update(block_policy="rl_semseg", block_target=0.3, block_train_interval=3, block_size=128, block_num_classes=19,
                block_cuda_graphs=True, block_policy_fused=os.environ.get("BLOCKCOPY_POLICY_FUSED", "1") != "0")
model = blockcopy.BlockCopyModel(build_swiftnet_rn18(), settings).eval().cuda().half()
model.policy.net = model.policy.net.float().train()
clip = synthetic_clip(30, 1024, 2048, seed=2, dtype=torch.float16, device="cuda")
timings.set_level(3)  # policy regions only: deeper levels synchronise inside the captured frame
with torch.no_grad():
    for rep in range(3):
        model.reset_temporal()
        if rep == 2:
            timings.reset() if hasattr(timings, "reset") else None
            torch.cuda.synchronize(); t0 = time.perf_counter()
        for f in clip:
            out = model(f)
            timings.add_cnt(1)
        torch.cuda.synchronize()
print("clip of 30 frames:", (time.perf_counter() - t0) * 1e3 / 30, "ms/frame (with timing syncs if level > 0)")
print(timings)
