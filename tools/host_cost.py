"""Host (Python + launch) cost per frame vs device time per frame of the benchmarked loop: the frames are issued
without synchronising, so the issue loop's wall clock is the host cost as long as the launch queue does not fill.
usage: python tools/host_cost.py [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from consumers.clips import synthetic_clip  # noqa: E402

sys.argv = [sys.argv[0]]
args = bench.parse_args()
dev = torch.device("cuda", 0)
n = 120
m = bench.build_model(args, dev)
clip = [f.to(dev) for f in synthetic_clip(30, 1024, 2048, seed=0, dtype=torch.float16)]
bench.run_frames([m], [clip], 0, 60, 30)
torch.cuda.synchronize()


def steady(count):  # no clip boundary inside: steady frames only (40 blocks)
    with torch.no_grad():
        for t in range(count):
            m(clip[1 + t % 29])


for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steady(n)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"issue {1e6 * (t1 - t0) / n:.1f} us/frame   issue+drain {1e6 * (t2 - t0) / n:.1f} us/frame")
# the graph replay alone
g = next(v for k, v in m._graphs.graphs.items() if k[0] == 40)[0]
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f"graph replay alone {a.elapsed_time(b) * 10:.1f} us/frame")
import cProfile, pstats  # noqa: E402
pr = cProfile.Profile()
pr.enable()
steady(n)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

# ---- the end-to-end pipeline (bench.HostPipeline): host issue cost vs wall clock per frame
host_clip = synthetic_clip(30, 1024, 2048, seed=0, dtype=torch.float16)
for u8 in (True, False):
    pipe = bench.HostPipeline([m], [host_clip], dev, u8=u8)
    pipe.run(0, 30, 30)
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        pipe.run(1, 29, 30)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"pipeline u8={u8}: issue {1e6 * (t1 - t0) / 29:.1f} us/frame   issue+drain {1e6 * (t2 - t0) / 29:.1f} us/frame")
    del pipe
