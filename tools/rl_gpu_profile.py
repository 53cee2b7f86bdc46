"""GPU time per kernel over 60 steady rl_semseg frames (torch profiler, kernels inside CUDA graphs included)."""
import random
import sys

import torch

sys.argv = ["bench.py"]
sys.path.insert(0, ".")
sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
import bench  # noqa: E402
from consumers.clips import synthetic_clip  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

args = bench.parse_args()
dev = torch.device("cuda", 0)
H, W, L = args.height, args.width, args.clip_length
random.seed(0)
torch.manual_seed(0)
m = bench.build_model(args, dev, policy="rl_semseg")
clip = [f.to(dev) for f in synthetic_clip(L, H, W, seed=3, dtype=torch.float16)]
bench.run_frames([m], [clip], 0, 4 * L, L)
torch.cuda.synchronize()
n = 60
with profile(activities=[ProfilerActivity.CUDA]) as prof, torch.no_grad():
    for t in range(n):
        k = bench._advance([m], t, L)
        m(clip[k])
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print("GPU time per frame: %.1f us over %d kernel launches per frame" % (tot / n, sum(r[1] for r in rows) / n))
for k, c, t in rows[:40]:
    print("  %7.1f us/frame  x%-5.1f  %s" % (t / n, c / n, k[:120]))
