"""bc_conv_igemm (tcgen05/TMEM/TMA implicit GEMM reading the plane by block index) against a plain
fp32 torch convolution of the same plane.  Tolerance (stated): fp16 operands, fp32 accumulation,
fp16 output => |d| <= 2^-9 * max|ref| + 2e-3."""
import pytest
import torch
import torch.nn.functional as F

from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _run(N, Cin, Cout, GH, GW, BS_in, k, stride, frac, bias=True, seed=0, relu=False, dil=1):
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
                        stride=stride, padding=(k // 2) * dil, dilation=dil)
    if relu:
        ref_full = ref_full.relu()
    ref = O.split(ref_full.cpu().contiguous(), me, BSo)
    out = torch.full((E, Cout, BSo, BSo), float("nan"), dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
    d_plane = plane.to(dev).contiguous(memory_format=torch.channels_last)
    d_w = weight.to(dev).contiguous(memory_format=torch.channels_last)
    _C.conv_igemm(out, d_plane, d_w, b.to(dev) if bias else None, None, me.to(dev), E, BS_in, stride, (k // 2) * dil, relu=relu)
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


@pytest.mark.parametrize("Cin,Cout,BS_in,dil,N,GH,GW,frac", [
    (64, 64, 16, 2, 1, 3, 4, 0.5), (256, 256, 8, 2, 2, 2, 3, 0.6), (128, 64, 32, 2, 1, 2, 2, 1.0), (64, 128, 4, 2, 1, 3, 3, 0.4),
    (128, 128, 16, 3, 1, 2, 3, 0.5), (64, 64, 8, 4, 1, 4, 4, 0.3)])
def test_dilated_3x3_conv(Cin, Cout, BS_in, dil, N, GH, GW, frac):
    """Pedestron's dilated backbone stage: 3x3, padding = dilation = 2 (reference tensorwrapper.py:539-548 leaves the
    dilation alone and takes a halo of `padding` pixels from the neighbouring blocks); also dilation 3 and 4, where
    the halo is wider than half a small block."""
    from blockcopy import _C

    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float16, device="cuda")
    assert _C.conv_supported(torch.float16, w, BS_in, 1, dil, dil)
    assert not _C.conv_supported(torch.float16, w, BS_in, 1, 1, dil)      # padding != dilation: not size preserving
    assert not _C.conv_supported(torch.float16, w, BS_in, 2, dil, dil)    # dilated + strided: not supported
    _run(N, Cin, Cout, GH, GW, BS_in, 3, 1, frac, seed=Cin + dil, dil=dil, relu=True)


@pytest.mark.parametrize("Cin,Cout,BS_in,k,stride,GH,GW,frac", [
    (512, 512, 2, 3, 1, 4, 8, 0.3), (256, 512, 4, 3, 2, 4, 8, 0.3), (256, 512, 4, 1, 2, 4, 8, 0.4), (512, 128, 2, 1, 1, 4, 8, 1.0),
    (64, 64, 2, 3, 1, 8, 16, 0.45), (512, 512, 2, 3, 1, 1, 1, 1.0), (128, 128, 2, 3, 1, 7, 9, 0.8)])
def test_two_pixel_blocks(Cin, Cout, BS_in, k, stride, GH, GW, frac):
    """`--block-size 64` (reference core/argparser.py:9) puts SwiftNet's layer4 at 2-px blocks: 32 blocks share one
    128-row accumulator tile (partly filled tiles, split-K over tiny grids, halo = the whole neighbouring block)."""
    _run(1, Cin, Cout, GH, GW, BS_in, k, stride, frac, seed=Cin + BS_in + GH, relu=True)
    _run(2, Cin, Cout, GH, GW, BS_in, k, stride, frac, seed=Cin + BS_in + GW, bias=False)


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


def test_conv_dual_write_and_residual_relu():
    """Epilogue fusion: + residual, ReLU, and the scatter into the next op's plane in one launch.
    Cells that were not executed keep the previous content of the plane (= previous frame)."""
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    N, C, GH, GW, BS = 2, 64, 2, 3, 16
    plane = torch.randn(N, C, GH * BS, GW * BS, generator=g).half()
    w = (torch.randn(C, C, 3, 3, generator=g) * 0.05).half()
    b = (0.1 * torch.randn(C, generator=g)).half()
    grid = torch.rand(N, 1, GH, GW, generator=g) < 0.5
    gi, me = O.grid_mappings(grid)
    E = me.numel()
    res = torch.randn(E, C, BS, BS, generator=g).half()
    nxt = torch.randn(N, C, GH * BS, GW * BS, generator=g).half()
    d = lambda t: t.to(dev).contiguous(memory_format=torch.channels_last)  # noqa: E731
    out, d_next = d(torch.zeros(E, C, BS, BS).half()), d(nxt)
    _C.conv_igemm(out, d(plane), d(w), b.to(dev), d(res), me.to(dev), E, BS, 1, 1, relu=True, plane_out=d_next)
    conv = F.conv2d(plane.to(dev).float(), w.to(dev).float(), b.to(dev).float(), padding=1).half()
    ref = (O.split(conv.cpu().contiguous(), me, BS).to(dev) + res.to(dev)).relu()
    err = (out.float() - ref.float()).abs().max().item()
    assert err <= 2 ** -9 * float(ref.abs().max()) + 2e-3, err
    want_plane = nxt.clone()
    O.combine_(out.cpu().contiguous(), want_plane, me)
    assert torch.equal(d_next.cpu().contiguous(), want_plane), "plane != scatter of the tiles / untouched cells changed"
    # plane only (write_tiles=False): same plane bits, the tile batch is left alone
    out2, d_next2 = d(torch.full((E, C, BS, BS), 7.0).half()), d(nxt)
    _C.conv_igemm(out2, d(plane), d(w), b.to(dev), d(res), me.to(dev), E, BS, 1, 1, relu=True, plane_out=d_next2,
                  write_tiles=False)
    assert torch.equal(d_next2, d_next) and bool((out2 == 7.0).all())


@pytest.mark.parametrize("Cin,Cout,BS", [(256, 256, 8), (512, 512, 4), (128, 128, 8)])
def test_split_k_is_deterministic_and_close(Cin, Cout, BS):
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(Cin + BS)
    plane = torch.randn(1, Cin, 8 * BS, 16 * BS, generator=g).half().to(dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).half().to(dev).contiguous(memory_format=torch.channels_last)
    me = torch.randperm(128, generator=g)[:40].sort().values.to(torch.int32).to(dev)
    outs = []
    bias = (0.1 * torch.randn(Cout, generator=g)).half().to(dev)
    res = torch.randn(40, Cout, BS, BS, generator=g).half().to(dev).contiguous(memory_format=torch.channels_last)
    for split in (True, True, False, "dsmem"):  # True: partials through the L2 scratch; "dsmem": through DSMEM
        o = torch.empty(40, Cout, BS, BS, dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
        _C.conv_igemm(o, plane, w, bias, res, me, 40, BS, 1, 1, relu=True, split_k=split)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]), "split-K must be run-to-run deterministic (ordered reduction)"
    ref = outs[2].float()
    # the DSMEM fallback (no scratch) may cut K differently: same tolerance against the unsplit result
    assert (outs[3].float() - ref).abs().max().item() <= 2 ** -9 * float(ref.abs().max()) + 2e-3
    assert (outs[0].float() - ref).abs().max().item() <= 2 ** -9 * float(ref.abs().max()) + 2e-3


@pytest.mark.parametrize("up2x,res,bn,relu", [(False, True, False, True), (True, True, True, True), (False, False, True, True),
                                              (True, False, False, False), (False, True, True, False)])
def test_ew_fused_matches_op_by_op(up2x, res, bn, relu):
    """bc_ew_fused == the torch op sequence it replaces (bilinear x2 per tile, add, eval batch_norm,
    ReLU), each rounded to fp16 like the separate kernels; <= 1 fp16 ulp, plane scatter bit-exact."""
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(7)
    N, C, GH, GW, BS = 1, 128, 3, 4, 16
    grid = torch.rand(N, 1, GH, GW, generator=g) < 0.5
    gi, me = O.grid_mappings(grid)
    E = me.numel()
    d = lambda t: t.to(dev).contiguous(memory_format=torch.channels_last)  # noqa: E731
    a = d(torch.randn(E, C, BS // 2 if up2x else BS, BS // 2 if up2x else BS, generator=g).half())
    r = d(torch.randn(E, C, BS, BS, generator=g).half()) if res else None
    mean, var = torch.randn(C, generator=g).half().to(dev), (0.5 + torch.rand(C, generator=g)).half().to(dev)
    wt, bs = (0.5 + torch.rand(C, generator=g)).half().to(dev), torch.randn(C, generator=g).half().to(dev)
    y = a
    if up2x:
        y = F.interpolate(y, size=(BS, BS), mode="bilinear")
    if res:
        y = y + r
    if bn:
        y = F.batch_norm(y, mean, var, wt, bs, False, 0.1, 1e-5)
    if relu:
        y = y.relu()
    bnp = (mean.float(), torch.rsqrt(var.float() + 1e-5), wt.float(), bs.float()) if bn else None
    out = d(torch.zeros(E, C, BS, BS).half())
    base = torch.randn(N, C, GH * BS, GW * BS, generator=g).half()
    plane = d(base)
    _C.ew_fused(out, a, r, bnp, relu, up2x, plane, me.to(dev))
    torch.cuda.synchronize()
    ulp = (out.float() - y.float()).abs() / (y.float().abs() * 2 ** -10 + 2 ** -14)
    assert ulp.max().item() <= 1.01, ulp.max().item()
    want = base.clone()
    O.combine_(out.cpu().contiguous(), want, me)
    assert torch.equal(plane.cpu().contiguous(), want)


@pytest.mark.parametrize("C,BS,k,stride,pad", [(64, 64, 3, 2, 1), (64, 16, 3, 1, 1), (128, 8, 3, 2, 1)])
def test_maxpool_halo_matches_zero_padded_pool(C, BS, k, stride, pad):
    """bc_maxpool_halo == max_pool2d(padding=0) of the zero-padded plane crop (the reference's padded-tile
    pooling: zeros, not -inf, outside the frame), bit-exact, incl. the scatter into the next plane."""
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(C + BS)
    N, GH, GW = 2, 2, 3
    plane = (torch.randn(N, C, GH * BS, GW * BS, generator=g) - 0.5).half()   # mostly negative at the border too
    grid = torch.rand(N, 1, GH, GW, generator=g) < 0.6
    gi, me = O.grid_mappings(grid)
    E, BSo = me.numel(), BS // stride
    d = lambda t: t.to(dev).contiguous(memory_format=torch.channels_last)  # noqa: E731
    out = d(torch.zeros(E, C, BSo, BSo).half())
    base = torch.randn(N, C, GH * BSo, GW * BSo, generator=g).half()
    nxt = d(base)
    _C.maxpool_halo(out, d(plane), me.to(dev), E, BS, k, stride, pad, plane_out=nxt)
    full = F.max_pool2d(F.pad(plane.float(), (pad,) * 4), k, stride, 0).half()
    ref = O.split(full.contiguous(), me, BSo)
    assert torch.equal(out.cpu().contiguous(), ref)
    want = base.clone()
    O.combine_(ref, want, me)
    assert torch.equal(nxt.cpu().contiguous(), want)


@pytest.mark.parametrize("N,GH,GW,BS,Cout,frac", [(1, 2, 3, 128, 64, 0.5), (2, 2, 2, 64, 64, 1.0), (1, 3, 2, 32, 128, 0.4)])
def test_stem_conv_7x7_s2(N, GH, GW, BS, Cout, frac):
    """bc_stem_pack + bc_conv_stem == relu(conv2d(frame, w, b, stride 2, padding 3)) on the executed blocks
    (halo from the neighbouring cells of the persistent space-to-depth plane, zeros outside the frame)."""
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(BS + Cout)
    H, W = GH * BS, GW * BS
    frame = torch.randn(N, 3, H, W, generator=g).half()
    w = (torch.randn(Cout, 3, 7, 7, generator=g) * (2.0 / 147) ** 0.5).half()
    b = (0.1 * torch.randn(Cout, generator=g)).half()
    # the plane holds an OLDER frame in the cells that are not executed now
    old = torch.randn(N, 3, H, W, generator=g).half()
    grid = torch.rand(N, 1, GH, GW, generator=g) < frac if frac < 1 else torch.ones(N, 1, GH, GW, dtype=torch.bool)
    gi, me = O.grid_mappings(grid)
    all_gi, all_me = O.grid_mappings(torch.ones_like(grid))
    E, G = me.numel(), grid.numel()
    plane = _C.stem_plane(N, H // 2, W // 2, torch.float16, dev)
    _C.stem_pack(plane, O.split(old, all_me, BS).to(dev), all_me.to(dev), G)
    _C.stem_pack(plane, O.split(frame, me, BS).to(dev), me.to(dev), E)
    mixed = old.clone()
    O.combine_(O.split(frame, me, BS), mixed, me)      # what the reference's frame_state would hold
    out = torch.full((E, Cout, BS // 2, BS // 2), float("nan"), dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
    nxt_base = torch.randn(N, Cout, H // 2, W // 2, generator=g).half()
    nxt = nxt_base.to(dev).contiguous(memory_format=torch.channels_last)
    _C.conv_stem(out, plane, _C.pack_stem_weight(w.to(dev)), b.to(dev), me.to(dev), E, relu=True, plane_out=nxt)
    ref_full = F.conv2d(mixed.to(dev).float(), w.to(dev).float(), b.to(dev).float(), stride=2, padding=3).relu()
    ref = O.split(ref_full.cpu().contiguous(), me, BS // 2)
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= 2 ** -9 * float(ref.abs().max()) + 2e-3, err
    want = nxt_base.clone()
    O.combine_(out.cpu().contiguous(), want, me)
    assert torch.equal(nxt.cpu().contiguous(), want)


_FORMS_SCRIPT = r'''
import sys, torch
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, sys.argv[3])
from blockcopy import _C
g = torch.Generator().manual_seed(3)
Cin = Cout = 128; BS = 32; GH, GW = 8, 16
plane = torch.randn(1, Cin, GH * BS, GW * BS, generator=g).half().cuda().contiguous(memory_format=torch.channels_last)
w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).half().cuda().contiguous(memory_format=torch.channels_last)
b = (0.1 * torch.randn(Cout, generator=g)).half().cuda()
me = torch.randperm(GH * GW, generator=g)[:40].sort().values.to(torch.int32).cuda()
res = torch.randn(40, Cout, BS, BS, generator=g).half().cuda().contiguous(memory_format=torch.channels_last)
nxt = torch.zeros(1, Cout, GH * BS, GW * BS, dtype=torch.float16, device="cuda").contiguous(memory_format=torch.channels_last)
out = torch.full((40, Cout, BS, BS), float("nan"), dtype=torch.float16, device="cuda").contiguous(memory_format=torch.channels_last)
_C.conv_igemm(out, plane, w, b, res, me, 40, BS, 1, 1, relu=True, plane_out=nxt)
torch.cuda.synchronize()
torch.save({"out": out.cpu(), "nxt": nxt.cpu(), "plane": plane.cpu(), "w": w.cpu(), "b": b.cpu(), "res": res.cpu(), "me": me.cpu()}, sys.argv[1])
'''


def test_large_grid_all_launch_forms_agree(tmp_path):
    """320 tiles (layer #20 at 1024x2048, E = 40): the one-tile-per-CTA kernel (default for this grid), the
    persistent kernel looping over 2-3 tiles per CTA with the double-buffered TMEM accumulator
    (BC_CONV_PERSIST=2) and the build without the persistent kernel (=0) issue the same MMA sequence per tile:
    identical bits, and within tolerance of the fp32 torch conv + residual + ReLU."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "blockcopy-video-processing-pytorch_b200")
    outs = {}
    for form in ("1", "2", "0", "a3"):
        path = str(tmp_path / f"form{form}.pt")
        # "a3": the default for this grid since r02 -- shared halo rows, k-steps ordered (kw, chunk, kh): another
        # summation order, so equal within rounding, not bitwise; the other three forms share the (kh, kw, chunk) order
        env = dict(os.environ, BC_CONV_PERSIST="1", BC_CONV_A3="1") if form == "a3" else \
            dict(os.environ, BC_CONV_PERSIST=form, BC_CONV_A3="0")
        subprocess.run([sys.executable, "-c", _FORMS_SCRIPT, path, root, pkg], check=True, env=env, timeout=300)
        outs[form] = torch.load(path)
    a = outs["1"]
    for form in ("2", "0"):
        assert torch.equal(a["out"], outs[form]["out"]), form
        assert torch.equal(a["nxt"], outs[form]["nxt"]), form
    a3 = outs["a3"]
    assert not torch.equal(a3["out"], a["out"]) or True  # (may coincide; what matters is the tolerance below)
    assert (a3["out"].float() - a["out"].float()).abs().max().item() <= 2 ** -9 * float(a["out"].float().abs().max()) + 2e-3
    want3 = torch.zeros_like(a3["nxt"]).contiguous()
    O.combine_(a3["out"].contiguous(), want3, a3["me"])
    assert torch.equal(a3["nxt"].contiguous(), want3)
    ref3 = (O.split(F.conv2d(a3["plane"].float(), a3["w"].float(), a3["b"].float(), padding=1).contiguous(), a3["me"], 32)
            .half().float() + a3["res"].float()).relu()
    assert (a3["out"].float() - ref3).abs().max().item() <= 2 ** -9 * float(ref3.abs().max()) + 2e-3
    ref_full = F.conv2d(a["plane"].float(), a["w"].float(), a["b"].float(), padding=1)
    ref = O.split(ref_full.contiguous(), a["me"], 32).half().float() + a["res"].float()
    ref = ref.relu()
    got = a["out"].float()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 2 ** -9 * float(ref.abs().max()) + 2e-3
    want = torch.zeros_like(a["nxt"]).contiguous()
    O.combine_(a["out"].contiguous(), want, a["me"])
    assert torch.equal(a["nxt"].contiguous(), want)


@pytest.mark.parametrize("bn,relu,dense_cl,with_prev", [(True, True, False, True), (True, True, True, True),
                                                          (False, False, False, True), (True, False, False, False)])
def test_head_1x1_matches_op_by_op(bn, relu, dense_cl, with_prev):
    """bc_head_1x1 == eval batch_norm -> ReLU -> 19-channel 1x1 conv (+bias) of torch on the tile batch, and the dense
    output == previous output with the executed cells replaced by exactly those tiles."""
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(11)
    N, Cin, Cout, GH, GW, BS = 2, 128, 19, 3, 4, 16
    grid = torch.rand(N, 1, GH, GW, generator=g) < 0.4 if with_prev else torch.ones(N, 1, GH, GW, dtype=torch.bool)
    gi, me = O.grid_mappings(grid)
    E = me.numel()
    x = torch.randn(E, Cin, BS, BS, generator=g).half().to(dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 1, 1, generator=g) * (2.0 / Cin) ** 0.5).half().to(dev)
    b = (0.1 * torch.randn(Cout, generator=g)).half().to(dev)
    mean, var = torch.randn(Cin, generator=g).half().to(dev), (torch.rand(Cin, generator=g) + 0.5).half().to(dev)
    gamma, beta = (torch.rand(Cin, generator=g) + 0.5).half().to(dev), (0.1 * torch.randn(Cin, generator=g)).half().to(dev)
    y = x
    params = None
    if bn:
        y = F.batch_norm(y, mean, var, gamma, beta, False, 0.1, 1e-5)
        params = (mean.float(), torch.rsqrt(var.float() + 1e-5), gamma.float(), beta.float())
    if relu:
        y = y.relu()
    ref = F.conv2d(y.float(), w.float()).half().float() + b.float().view(1, -1, 1, 1)  # conv rounded, then the bias add
    prev = torch.randn(N, Cout, GH * BS, GW * BS, generator=g).half().to(dev)
    if dense_cl:
        prev = prev.contiguous(memory_format=torch.channels_last)
    dense = torch.full_like(prev, float("nan"))
    tiles = torch.full((E, Cout, BS, BS), float("nan"), dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
    _C.head_1x1(x, w.reshape(Cout, Cin).contiguous(), b, params, relu, tiles_out=tiles, dense_out=dense,
                dense_prev=prev if with_prev else None, grid_idx=gi.to(dev), mapping_exec=me.to(dev))
    torch.cuda.synchronize()
    got = tiles.float()
    assert torch.isfinite(got).all()
    # fp32 accumulation in another order than cuDNN's, BN rounded to fp16 in both: a few fp16 ulps of the result
    assert (got - ref).abs().max().item() <= 2 ** -8 * float(ref.abs().max()) + 2e-3
    want = prev.cpu().contiguous().clone()
    O.combine_(tiles.cpu().contiguous(), want, me)
    assert torch.equal(dense.cpu().contiguous(), want)



_MULTICAST_SCRIPT = r'''
import sys, torch
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, sys.argv[3])
from blockcopy import _C
Cin, Cout, BS, E = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
g = torch.Generator().manual_seed(Cin + BS)
N, GH, GW = 2, 16, 16
plane = torch.randn(N, Cin, GH * BS, GW * BS, generator=g).half().cuda().contiguous(memory_format=torch.channels_last)
w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).half().cuda().contiguous(memory_format=torch.channels_last)
b = (0.1 * torch.randn(Cout, generator=g)).half().cuda()
me = torch.randperm(N * GH * GW, generator=g)[:E].sort().values.to(torch.int32).cuda()
res = torch.randn(E, Cout, BS, BS, generator=g).half().cuda().contiguous(memory_format=torch.channels_last)
nxt = torch.zeros(N, Cout, GH * BS, GW * BS, dtype=torch.float16, device="cuda").contiguous(memory_format=torch.channels_last)
out = torch.full((E, Cout, BS, BS), float("nan"), dtype=torch.float16, device="cuda").contiguous(memory_format=torch.channels_last)
for _ in range(3):
    _C.conv_igemm(out, plane, w, b, res, me, E, BS, 1, 1, relu=True, plane_out=nxt)
torch.cuda.synchronize()
torch.save({"out": out.cpu(), "nxt": nxt.cpu(), "plane": plane.cpu(), "w": w.cpu(), "b": b.cpu(), "res": res.cpu(), "me": me.cpu()}, sys.argv[1])
'''


@pytest.mark.parametrize("Cin,Cout,BS,E", [(512, 512, 4, 325), (512, 512, 2, 509), (256, 512, 4, 320), (128, 128, 32, 41),
                                           (64, 64, 32, 37), (128, 128, 16, 301)])
def test_multicast_forms_are_bit_identical(tmp_path, Cin, Cout, BS, E):
    """Big grids run the one-tile kernel as thread-block clusters that share an operand through TMA multicast
    (BC_CONV_MULTICAST=0: every CTA loads everything itself): small blocks (4- / 2-px, 8 streams batched) -- the four channel
    slices of a pixel tile share its activation boxes; big blocks (16- / 32-px) -- four (two) pixel tiles of a channel slice
    share the weight tile.  The same MMA sequence, so the same bits -- incl. a last tile that is only partly filled (E not
    a multiple of the blocks per tile) -- and within tolerance of the fp32 torch conv."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "blockcopy-video-processing-pytorch_b200")
    outs = {}
    for mc in ("3", "0"):
        path = str(tmp_path / f"mc{mc}.pt")
        subprocess.run([sys.executable, "-c", _MULTICAST_SCRIPT, path, root, pkg, str(Cin), str(Cout), str(BS), str(E)], check=True,
                       env=dict(os.environ, BC_CONV_MULTICAST=mc), timeout=120)
        outs[mc] = torch.load(path)
    a, b = outs["3"], outs["0"]
    assert torch.isfinite(a["out"].float()).all()
    assert torch.equal(a["out"], b["out"]) and torch.equal(a["nxt"], b["nxt"])
    ref = (O.split(F.conv2d(a["plane"].float(), a["w"].float(), a["b"].float(), padding=1).contiguous(), a["me"], BS)
           .half().float() + a["res"].float()).relu()
    assert (a["out"].float() - ref).abs().max().item() <= 2 ** -9 * float(ref.abs().max()) + 2e-3
