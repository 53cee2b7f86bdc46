"""Where a policy training step spends its GPU time: torch autograd over cuDNN (graph replays) vs policy/fused_train.py;
per-kernel table of one eager native forward + backward."""
import sys

import torch

sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
sys.path.insert(0, "tests")
from test_gpu_policy_train import _policy_net  # noqa: E402

from blockcopy.policy.fused_train import FusedPolicyTrainer  # noqa: E402

N, H, W = 1, 256, 512
dev = torch.device("cuda")
x = torch.randn(N, 26, H, W, device=dev)
R = torch.randn(N, 1, H // 32, W // 32, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


# torch path, graphed (what PolicyNet._trunk_forward does with use_cuda_graphs)
net = _policy_net(0).to(memory_format=torch.channels_last)
net.use_cuda_graphs = True
net.channels_last = True


def torch_step():
    for q in net.parameters():
        q.grad = None
    (net._trunk_forward(x) * R).mean().backward()


print("torch graphed fwd+bwd: %.1f us" % timed(torch_step))
net2 = _policy_net(0)
tr = FusedPolicyTrainer(net2)
tr.use_cuda_graph = True
fill = lambda x16: x16[:, :26].copy_(x)  # noqa: E731


def ours_step():
    (tr.run_train(fill, (N, 26, H, W), dev) * R).mean().backward()


print("native graphed fwd+bwd: %.1f us" % timed(ours_step))
print("  forward graph replay: %.1f us" % timed(lambda: tr._fwd_graph[1].replay()))
print("  backward graph replay: %.1f us" % timed(lambda: tr._bwd_graph[0].replay()))
tr.use_cuda_graph = False
tr._eager = True
from torch.profiler import ProfilerActivity, profile  # noqa: E402

ours_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ours_step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print("eager native step: %d kernels, %.1f us of GPU time" % (sum(r[1] for r in rows), tot))
for k, c, t in rows[:22]:
    print("  %7.1f us  x%-3d  %s" % (t, c, k[:110]))

# host side of the graphed native step
import cProfile  # noqa: E402
import pstats  # noqa: E402
import time  # noqa: E402

tr.use_cuda_graph = True
for _ in range(3):
    ours_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    ours_step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue time per native graphed step: %.1f us" % ((t1 - t0) * 1e6 / 20))
t0 = time.perf_counter()
for _ in range(20):
    torch_step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue time per torch graphed step: %.1f us" % ((t1 - t0) * 1e6 / 20))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    ours_step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
