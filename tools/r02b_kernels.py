"""One launch of each kernel added in round 2b at its representative size (for `ncu --set full`)."""
import sys

import torch

sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
from blockcopy import _C  # noqa: E402
from blockcopy.core.frame import CITYSCAPES_MEAN as M, CITYSCAPES_STD as S  # noqa: E402

dev = torch.device("cuda")
cl = lambda t: t.contiguous(memory_format=torch.channels_last)  # noqa: E731
for rep in range(2):
    # policy backward at the 256x512 layers (64 padded channels)
    z = cl(torch.randn(1, 64, 256, 512, device=dev).half())
    g = cl(torch.randn(1, 64, 256, 512, device=dev).half())
    out = cl(torch.randn(1, 64, 256, 512, device=dev).half())
    mean, invstd, gamma = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.ones(64, device=dev)
    ws = torch.zeros(_C.BN_STATS_WORKSPACE, dtype=torch.uint8, device=dev)
    sums = torch.empty(2, 64, device=dev)
    _C.bn_bwd_reduce(sums, g, out, z, mean, invstd, ws)
    dz = torch.empty_like(z)
    _C.bn_bwd_apply(dz, None, g, out, z, mean, invstd, gamma, sums)
    grad = torch.empty(32, 32, 3, 3, device=dev)
    wws = torch.empty(_C.WGRAD_WORKSPACE, dtype=torch.uint8, device=dev)
    _C.conv_wgrad(grad, dz, z, 1, None, wws)
    grad2 = torch.empty(128, 128, 3, 3, device=dev)
    x2 = cl(torch.randn(1, 128, 64, 128, device=dev).half())
    _C.conv_wgrad(grad2, x2, x2, 1, None, wws)
    # CSP ops, E = 40
    t = cl(torch.randn(40, 256, 32, 32, device=dev).half())
    gws = torch.zeros(_C.GN_STATS_WORKSPACE, dtype=torch.uint8, device=dev)
    st = torch.empty(2, 256, device=dev)
    _C.gn_stats(t, 32, 1e-5, st[0], st[1], gws)
    ph = cl(torch.randn(40, 4096, 8, 8, device=dev).half())
    up = cl(torch.empty(40, 256, 32, 32, dtype=torch.float16, device=dev))
    _C.depth_to_space(up, ph, 4)
    # first gather from the uint8 frame
    u8 = torch.randint(0, 256, (1, 1024, 2048, 3), dtype=torch.uint8, device=dev)
    cells = torch.randperm(128)[:40].sort().values.to(torch.int32).to(dev)
    tiles = torch.empty(40, 3, 128, 128, dtype=torch.float16, device=dev)
    _C.blocks_from_u8(tiles, u8, M, S, cells, 40)
    torch.cuda.synchronize()
print("ok")
