"""Fused pyramid pooling (bc_spp_* + bc_ew_fused + bc_conv_igemm) against the module's own dense torch forward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,H,W", [(1, 32, 64), (2, 16, 32), (1, 8, 16)])
def test_fused_spp_matches_module(N, H, W):
    from blockcopy.core.fused_spp import try_fused_spp
    from consumers.clips import deterministic_init_
    from consumers.swiftnet_rn18 import SpatialPyramidPooling

    spp = deterministic_init_(SpatialPyramidPooling(512, bt_size=128, level_size=42, out_size=128), seed=3)
    spp = spp.eval().cuda().half().to(memory_format=torch.channels_last)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, 512, H, W, generator=g).half().cuda().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        ref = spp(x)                    # plain tensor in -> the module's own forward, op by op
        got = try_fused_spp(spp, x)
    assert got is not None and got.shape == ref.shape and got.dtype == torch.float16
    tol = 2 ** -7 * float(ref.float().abs().max()) + 1e-3
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= tol, (err, tol)
    # weights changed in place -> the cached plan must notice
    with torch.no_grad():
        spp.spp[0].conv.weight.mul_(0.5)
        ref2, got2 = spp(x), try_fused_spp(spp, x)
    assert (got2.float() - ref2.float()).abs().max().item() <= 2 ** -7 * float(ref2.float().abs().max()) + 1e-3


def test_fused_spp_declines_what_it_does_not_cover():
    from blockcopy.core.fused_spp import try_fused_spp
    from consumers.swiftnet_rn18 import SpatialPyramidPooling

    spp = SpatialPyramidPooling(512).eval().cuda().half()
    x = torch.randn(1, 512, 30, 64).half().cuda()
    assert try_fused_spp(spp, x) is None                      # 30 rows: grids do not divide
    assert try_fused_spp(spp.float(), x.float()) is None       # fp32
    assert try_fused_spp(torch.nn.Conv2d(512, 8, 1).cuda().half(), torch.randn(1, 512, 32, 64).half().cuda()) is None
