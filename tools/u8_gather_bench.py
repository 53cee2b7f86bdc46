"""First gather of a frame: bc_gather from the normalised fp16 frame vs bc_blocks_from_u8 from the uint8 frame
(1024x2048, 128-px blocks, E = 40), us per launch inside a CUDA graph of 100 launches over 4 rotating frames."""
import sys

import torch

sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
from blockcopy import _C  # noqa: E402
from blockcopy.core.frame import CITYSCAPES_MEAN as M, CITYSCAPES_STD as S  # noqa: E402

dev = torch.device("cuda")
H, W, BS, E = 1024, 2048, 128, 40
g = torch.Generator().manual_seed(0)
cells = torch.randperm(128, generator=g)[:E].sort().values.to(torch.int32).to(dev)
u8 = [torch.randint(0, 256, (1, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(4)]
f16 = [_C.frame_from_u8(u, M, S, torch.float16) for u in u8]
tiles = [torch.empty(E, 3, BS, BS, dtype=torch.float16, device=dev) for _ in range(4)]


def timed(fn, reps=100):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(reps):
            fn(i)
    gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    gr.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


print("bc_gather (fp16 frame -> tiles): %.2f us" % timed(lambda i: _C.gather(tiles[i % 4], f16[i % 4], cells, E)))
print("bc_blocks_from_u8 (u8 frame -> tiles): %.2f us" % timed(lambda i: _C.blocks_from_u8(tiles[i % 4], u8[i % 4], M, S, cells, E)))
print("bc_frame_from_u8 (whole frame): %.2f us" % timed(lambda i: _C.frame_from_u8(u8[i % 4], M, S, torch.float16, f16[i % 4])))
