"""Golden vectors for the driver-side steps, made with the calls the reference's driver makes --
torchvision's to_tensor + normalize (semantic_segmentation/lib/ext_transforms.py:317-372 wrap exactly these) and
F.interpolate(mode='bilinear') + max(dim=1) (test_swiftnet.py:196-197) -- on the CPU, in THIS container:

    python oracle/make_golden_io.py        ->  tests/golden/io_kat.pt

Small on purpose (a 24x40 frame pair, 19x8x12 logits): the fixture pins the oracle (tests/test_oracle_io.py) and
the kernels (tests/test_gpu_io.py) to outputs that neither of them produced.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
import torchvision.transforms.functional as TF
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEAN = (73.1584 / 255, 82.9090 / 255, 72.3924 / 255)   # lib/datasets/cityscapes_vid.py:29-30
STD = (44.9149 / 255, 46.1529 / 255, 45.3192 / 255)


def main():
    rng = np.random.default_rng(2024)
    u8 = rng.integers(0, 256, size=(2, 24, 40, 3), dtype=np.uint8)
    u8[0, 0, :, 0] = np.arange(40) * 6
    u8[1, 3, :16] = 255
    frames = torch.stack([TF.normalize(TF.to_tensor(Image.fromarray(f)), MEAN, STD) for f in u8])
    g = torch.Generator().manual_seed(2024)
    logits32 = 2 * torch.randn(2, 19, 8, 12, generator=g)
    logits16 = logits32.half()
    up32 = F.interpolate(logits32, size=(32, 48), mode="bilinear")
    # fp16 logits: the upsampled tensor the driver would hold is fp16; ties (rare) go to the lowest class index
    up16 = F.interpolate(logits16.float(), size=(32, 48), mode="bilinear").half()
    vals16 = up16.float().max(dim=1, keepdim=True).values
    out = {
        "mean": MEAN, "std": STD, "u8": torch.from_numpy(u8),
        "frames_fp32": frames, "frames_fp16": frames.half(),
        "logits32": logits32, "labels32": up32.max(dim=1)[1],
        "logits16": logits16, "labels16": (up16.float() == vals16).float().argmax(dim=1),
        "top2_gap32": (lambda t: t[:, 0] - t[:, 1])(up32.topk(2, dim=1).values),
        "made_with": f"torch {torch.__version__}, torchvision {sys.modules['torchvision'].__version__}",
    }
    path = os.path.join(ROOT, "tests", "golden", "io_kat.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
