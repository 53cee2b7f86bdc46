"""Policy trunk forward at 1024x2048 (features 1x26x256x512): torch (graphed, autograd-aware) vs policy/fused_net.py
(graph replay), GPU time per call; plus the largest logit difference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))
import torch
from blockcopy.policy.net import PolicyNet
from blockcopy.policy.fused_net import FusedPolicyTrunk

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = PolicyNet(block_size=128, task_num_classes=19).to(dev).train()
net.channels_last = True
net.to(memory_format=torch.channels_last)
net.use_cuda_graphs = True
x = torch.randn(1, 26, 256, 512, device=dev)
fused = FusedPolicyTrunk(net)


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


with torch.enable_grad():
    t_torch = timed(lambda: net._trunk_forward(x))
    want = net._trunk_forward(x).detach().clone()
t_fused = timed(lambda: fused(x, use_cuda_graph=True))
t_eager = timed(lambda: fused(x, use_cuda_graph=False), reps=10)
got = fused(x, use_cuda_graph=True)
print(f"torch graphed trunk forward: {t_torch:.1f} us | fused graph replay: {t_fused:.1f} us | fused eager: {t_eager:.1f} us")
print(f"max |logit diff| {float((got - want).abs().max()):.4f} of range {float(want.abs().max()):.3f}; "
      f"max |prob diff| {float((torch.sigmoid(got) - torch.sigmoid(want)).abs().max()):.4f}")
if "--list" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fused(x, use_cuda_graph=False)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    print(" | ".join(f"{e.name.split('(')[0].replace('void ', '').replace('bc::', '')[:22]} {e.device_time:.0f}" for e in evs))
