"""Host side of the end-to-end loop (bench.HostPipeline, uint8 in / label map out): enqueue time per frame and the
cProfile table of 300 steady frames.  If the enqueue time is close to the e2e frame time, e2e is bound by Python."""
import cProfile
import pstats
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
model = bench.build_model(args, dev)
host_clip = synthetic_clip(L, H, W, seed=0, batch=1, dtype=torch.float16)
clip = [f.to(dev) for f in host_clip]
bench.run_frames([model], [clip], 0, 3 * L, L, False)
pipe = bench.HostPipeline([model], [host_clip], dev, u8=True)
pipe.run(0, 2 * L, L)
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
pipe.run(0, n, L)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.1f us/frame, until the GPU is done %.1f us/frame" % ((t1 - t0) * 1e6 / n, (t2 - t0) * 1e6 / n))
torch.cuda.synchronize()
t0 = time.perf_counter()
bench.run_frames([model], [clip], 0, n, L, False)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("device-resident loop: host enqueue %.1f us/frame, until the GPU is done %.1f us/frame" % ((t1 - t0) * 1e6 / n, (t2 - t0) * 1e6 / n))
pr = cProfile.Profile()
pr.enable()
pipe.run(0, n, L)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(32)
