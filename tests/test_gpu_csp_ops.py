"""Pedestron / CSP op set on native kernels: GroupNorm over all executed blocks (bc_gn_stats +
bc_ew_fused, reference core/tensorwrapper.py:600-633) and per-block ConvTranspose2d (bc_conv_igemm over the tile batch +
bc_depth_to_space; a pass-through op in the reference, tensorwrapper.py:519-520), against fp32 torch on the same tiles."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("E,C,h,groups", [(5, 32, 8, 4), (40, 256, 32, 32), (3, 64, 4, 1), (7, 128, 16, 16), (1, 8, 2, 1)])
def test_gn_stats_match_float64_and_are_reproducible(E, C, h, groups):
    from blockcopy import _C

    g = torch.Generator(device="cuda").manual_seed(C + h)
    x = (1.3 * torch.randn(E, C, h, h, device="cuda", generator=g) + 0.4).half().contiguous(memory_format=torch.channels_last)
    assert _C.gn_supported(x, groups)
    ws = torch.zeros(_C.GN_STATS_WORKSPACE, dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(3):
        mean, invstd = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        _C.gn_stats(x, groups, 1e-5, mean, invstd, ws)
        outs.append((mean, invstd))
    # the reference's fold: one sample whose group statistics run over all tiles
    x64 = x.double().permute(1, 0, 2, 3).reshape(groups, -1)
    want_mean = x64.mean(dim=1).repeat_interleave(C // groups)
    want_inv = (1.0 / torch.sqrt(x64.var(dim=1, unbiased=False) + 1e-5)).repeat_interleave(C // groups)
    assert torch.allclose(outs[0][0].double(), want_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(outs[0][1].double(), want_inv, rtol=1e-5, atol=1e-6)
    for m, s in outs[1:]:
        assert torch.equal(m, outs[0][0]) and torch.equal(s, outs[0][1])
    assert int(ws[:4].view(torch.int32)) == 0


def _blocks(C, BS, GH=2, GW=3, frac=0.6, seed=0):
    import blockcopy

    g = torch.Generator().manual_seed(seed)
    img = torch.randn(1, C, GH * BS, GW * BS, generator=g).half().cuda()
    grid = torch.ones(1, 1, GH, GW, dtype=torch.bool)
    x = blockcopy.to_tensorwrapper(img)
    feats = x.process_temporal_features(None)
    x = x.to_blocks(grid.cuda())
    grid2 = torch.rand(1, 1, GH, GW, generator=g) < frac
    grid2[0, 0, 0, 0] = True
    y = blockcopy.to_tensorwrapper(img)
    y.process_temporal_features(feats)
    return y.to_blocks(grid2.cuda())


@pytest.mark.parametrize("C,BS,groups,affine,relu", [(32, 8, 4, True, True), (256, 16, 32, True, False), (64, 4, 8, False, True)])
def test_group_norm_on_blocks_uses_native_kernels(C, BS, groups, affine, relu):
    """F.group_norm on a block tensor == torch on the reference's fold (tiles -> one sample), within fp16 rounding; and
    the statistics + normalisation run as bc_gn_stats + bc_ew_fused (a deferred tensor comes back)."""
    b = _blocks(C, BS, seed=C)
    tiles = b.as_subclass(torch.Tensor).clone()
    E = tiles.shape[0]
    g = torch.Generator().manual_seed(1)
    w = (torch.rand(C, generator=g) + 0.5).half().cuda() if affine else None
    bias = (torch.randn(C, generator=g) * 0.2).half().cuda() if affine else None
    out = F.group_norm(b, groups, w, bias, 1e-5)
    assert out._pending is not None and out._pending.kind == "ew"
    if relu:
        out = F.relu(out, inplace=True)
    got = out.as_subclass(torch.Tensor).float() if out._pending is None else None
    if got is None:
        out._materialize()
        got = out.as_subclass(torch.Tensor).float()
    folded = tiles.float().permute(1, 0, 2, 3).reshape(1, C, E * BS, BS)
    want = F.group_norm(folded, groups, None if w is None else w.float(), None if bias is None else bias.float(), 1e-5)
    want = want.reshape(C, E, BS, BS).permute(1, 0, 2, 3)
    if relu:
        want = want.relu()
    assert (got - want).abs().max().item() <= 2 ** -9 * float(want.abs().max()) + 2e-3


@pytest.mark.parametrize("E,C,h,r", [(3, 16, 4, 2), (5, 64, 8, 4), (1, 8, 2, 2), (40, 256, 16, 2)])
def test_depth_to_space_is_bit_exact(E, C, h, r):
    from blockcopy import _C

    g = torch.Generator(device="cuda").manual_seed(E)
    x = torch.randn(E, r * r * C, h, h, device="cuda", generator=g).half().contiguous(memory_format=torch.channels_last)
    out = torch.empty(E, C, r * h, r * h, dtype=torch.float16, device="cuda").contiguous(memory_format=torch.channels_last)
    _C.depth_to_space(out, x, r)
    want = x.reshape(E, r, r, C, h, h).permute(0, 3, 4, 1, 5, 2).reshape(E, C, r * h, r * h)
    assert torch.equal(out, want)


@pytest.mark.parametrize("Cin,Cout,BS,k,s,p,bias", [(64, 16, 8, 4, 2, 1, True), (64, 16, 4, 4, 4, 0, True), (128, 64, 16, 4, 2, 1, False),
                                                    (256, 256, 8, 4, 4, 0, True), (64, 64, 2, 2, 2, 0, True)])
def test_conv_transpose_on_blocks_matches_fp32_torch(Cin, Cout, BS, k, s, p, bias):
    from blockcopy import _C

    b = _blocks(Cin, BS, seed=Cin + BS)
    tiles = b.as_subclass(torch.Tensor).clone()
    g = torch.Generator().manual_seed(2)
    w = (torch.randn(Cin, Cout, k, k, generator=g) * (2.0 / (Cin * 4)) ** 0.5).half().cuda()
    bb = (0.1 * torch.randn(Cout, generator=g)).half().cuda() if bias else None
    assert _C.deconv_supported(torch.float16, w, BS, s, p)
    out = F.conv_transpose2d(b, w, bb, stride=s, padding=p)
    assert out.is_blocks and tuple(out.shape) == (tiles.shape[0], Cout, s * BS, s * BS)
    want = F.conv_transpose2d(tiles.float(), w.float(), None if bb is None else bb.float(), stride=s, padding=p)
    got = out.as_subclass(torch.Tensor).float()
    assert (got - want).abs().max().item() <= 2 ** -9 * float(want.abs().max()) + 2e-3
    # outside the envelope: torch on the tile batch, same numbers within fp16
    out2 = F.conv_transpose2d(b, w, bb, stride=s, padding=p, output_padding=1 if s > 1 and k == 4 and p == 1 else 0)
    assert out2.is_blocks
