"""bc_conv_igemm (tcgen05/TMEM/TMA implicit GEMM reading the plane by block index) against a plain
fp32 torch convolution of the same plane.  Tolerance (stated): fp16 operands, fp32 accumulation,
fp16 output => |d| <= 2^-9 * max|ref| + 2e-3."""
import pytest
import torch
import torch.nn.functional as F

from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _run(N, Cin, Cout, GH, GW, BS_in, k, stride, frac, bias=True, seed=0, relu=False):
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(seed)
    H, W = GH * BS_in, GW * BS_in
    plane = torch.randn(N, Cin, H, W, generator=g).half()
    weight = (torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5).half()
    b = (0.1 * torch.randn(Cout, generator=g)).half() if bias else None
    grid = torch.rand(N, 1, GH, GW, generator=g) < frac if frac < 1 else torch.ones(N, 1, GH, GW, dtype=torch.bool)
    gi, me = O.grid_mappings(grid)
    E = me.numel()
    if E == 0:
        return
    BSo = BS_in // stride
    ref_full = F.conv2d(plane.to(dev).float(), weight.to(dev).float(), b.to(dev).float() if bias else None,
                        stride=stride, padding=k // 2)
    if relu:
        ref_full = ref_full.relu()
    ref = O.split(ref_full.cpu().contiguous(), me, BSo)
    out = torch.full((E, Cout, BSo, BSo), float("nan"), dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
    d_plane = plane.to(dev).contiguous(memory_format=torch.channels_last)
    d_w = weight.to(dev).contiguous(memory_format=torch.channels_last)
    _C.conv_igemm(out, d_plane, d_w, b.to(dev) if bias else None, None, me.to(dev), E, BS_in, stride, k // 2, relu=relu)
    torch.cuda.synchronize()
    got = out.float().cpu()
    assert torch.isfinite(got).all(), "unwritten or non-finite outputs"
    tol = 2 ** -9 * float(ref.abs().max()) + 2e-3
    err = (got - ref).abs().max().item()
    assert err <= tol, (err, tol)


# every 3x3 / 1x1 conv shape SwiftNet-RN18 issues on blocks, 128-px image blocks
@pytest.mark.parametrize("Cin,Cout,BS_in,k,stride", [
    (64, 64, 32, 3, 1), (64, 128, 32, 3, 2), (64, 128, 32, 1, 2), (128, 128, 16, 3, 1), (128, 256, 16, 3, 2),
    (128, 256, 16, 1, 2), (256, 256, 8, 3, 1), (256, 512, 8, 3, 2), (256, 512, 8, 1, 2), (512, 512, 4, 3, 1),
    (256, 128, 8, 1, 1), (128, 128, 16, 1, 1), (64, 128, 32, 1, 1), (128, 128, 8, 3, 1), (128, 128, 32, 3, 1),
])
def test_swiftnet_conv_shapes(Cin, Cout, BS_in, k, stride):
    _run(1, Cin, Cout, 3, 4, BS_in, k, stride, 0.4, seed=Cin + Cout + BS_in)


@pytest.mark.parametrize("N,GH,GW,frac", [(1, 2, 2, 1.0), (2, 2, 3, 0.5), (1, 1, 1, 1.0), (3, 3, 3, 0.2)])
def test_conv_grids_and_batches(N, GH, GW, frac):
    _run(N, 64, 64, GH, GW, 16, 3, 1, frac, seed=N * 10 + GH)
    _run(N, 128, 128, GH, GW, 8, 3, 1, frac, bias=False, seed=N * 10 + GW, relu=True)


def test_identity_mapping_1x1_on_tiles():
    """1x1 convs take the packed tile batch itself as the 'plane' (mapping NULL)."""
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    tiles = torch.randn(5, 128, 8, 8, generator=g).half()
    w = (torch.randn(64, 128, 1, 1, generator=g) * 0.1).half()
    out = torch.empty(5, 64, 8, 8, dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
    _C.conv_igemm(out, tiles.to(dev).contiguous(memory_format=torch.channels_last), w.to(dev), None, None, None, 5, 8, 1, 0)
    ref = F.conv2d(tiles.float(), w.float())
    assert (out.float().cpu() - ref).abs().max().item() <= 2 ** -9 * float(ref.abs().max()) + 2e-3


def test_unsupported_shapes_are_refused_not_miscomputed():
    from blockcopy import _C

    dev = "cuda"
    t = torch.zeros(1, 48, 8, 8, dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.zeros(64, 48, 3, 3, dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
    assert not _C.conv_supported(t.dtype, w, 8, 1, 1)
    out = torch.zeros(1, 64, 8, 8, dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
    with pytest.raises(_C.BlockCopyNativeError, match="multiple of 64"):
        _C.conv_igemm(out, t, w, None, None, None, 1, 8, 1, 1)
