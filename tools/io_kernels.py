"""A few launches of the driver-side kernels at BASELINE size, for `ncu --set full` (tools/gpu_prof_io.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))
import torch
from consumers.frame_io import FrameNormalizer, predict_labels

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
u8 = [torch.randint(0, 256, (1, 1024, 2048, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(4)]
lg = [torch.randn(1, 19, 256, 512, device=dev, generator=g).half() for _ in range(4)]
norm = FrameNormalizer()
for i in range(4):
    norm(u8[i])
    predict_labels(lg[i])
torch.cuda.synchronize()
print("ok")
