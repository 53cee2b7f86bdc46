"""Host side of rl_semseg frames (bench.bench_rl's loop): where the Python thread waits for the GPU (syncs) and what it
spends on enqueueing; cProfile of 180 steady frames."""
import cProfile
import pstats
import random
import sys
import time

import torch

sys.argv = ["bench.py"]
sys.path.insert(0, ".")
sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
import bench  # noqa: E402
from consumers.clips import synthetic_clip  # noqa: E402

args = bench.parse_args()
dev = torch.device("cuda", 0)
H, W, L = args.height, args.width, args.clip_length
random.seed(0)
torch.manual_seed(0)
m = bench.build_model(args, dev, policy="rl_semseg")
clip = [f.to(dev) for f in synthetic_clip(L, H, W, seed=3, dtype=torch.float16)]
bench.run_frames([m], [clip], 0, 4 * L, L)
torch.cuda.synchronize()


def loop(n):
    with torch.no_grad():
        for t in range(n):
            k = bench._advance([m], t, L)
            m(clip[k])


t0 = time.perf_counter()
loop(180)
torch.cuda.synchronize()
print("rl_semseg: %.1f us/frame wall" % ((time.perf_counter() - t0) * 1e6 / 180))
pr = cProfile.Profile()
pr.enable()
loop(180)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(26)
