import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))
import torch
import bench
from blockcopy import _C
dev = torch.device("cuda", 0)
peaks = bench.load_peaks()
split = os.environ.get("SPLIT", "1") == "1"
orig = _C.conv_igemm
_C.conv_igemm = lambda *a, **k: orig(*a, **dict(k, split_k=split))
r = bench.conv_microbench(dev, peaks)
print("SPLIT", split, "DEBUG", os.environ.get("BC_CONV_DEBUG"), {k: round(v["us"], 1) for k, v in r.items()})
