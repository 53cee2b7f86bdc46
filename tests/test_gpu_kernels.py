"""Parity of the sm_100a block kernels (through the C ABI) against the CPU oracle: bit-exact,
both layouts, both dtypes, SIMT and TMA paths, every (C, BS, p) SwiftNet-RN18 issues
(SURVEY.md 3.2), BASELINE config 2, edge cases (E=0, E=all, N=2, BS<=2, p=2,3, ragged grids)."""
import pytest
import torch

from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _bits(t):
    t = t.detach().cpu().contiguous()
    return t.view(torch.int16 if t.element_size() == 2 else torch.int32)


def _same(a, b):
    return torch.equal(_bits(a), _bits(b))


def _fmt(t, nhwc):
    return t.contiguous(memory_format=torch.channels_last) if nhwc else t.contiguous()


def _rand(shape, dtype, g):
    return torch.randn(shape, generator=g).to(dtype)


def _grid(N, GH, GW, frac, g):
    if frac >= 1:
        return torch.ones(N, 1, GH, GW, dtype=torch.bool)
    if frac <= 0:
        return torch.zeros(N, 1, GH, GW, dtype=torch.bool)
    return torch.rand(N, 1, GH, GW, generator=g) < frac


def _run_all(N, C, GH, GW, BS, pad, dtype, nhwc, frac, seed=0):
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator().manual_seed(seed)
    H, W = GH * BS, GW * BS
    plane = _rand((N, C, H, W), dtype, g)
    prevp = _rand((N, C, H, W), dtype, g)
    grid = _grid(N, GH, GW, frac, g)
    prev_grid = _grid(N, GH, GW, 0.5, g)
    G = grid.numel()

    # ---- index tensors ------------------------------------------------------------------------
    pgi, pme = O.grid_mappings(prev_grid)
    gi, me = O.grid_mappings(grid)
    ti = O.transfer_idx(grid, pgi)
    d_gi = torch.full((N, 1, GH, GW), 12345, dtype=torch.int32, device=dev)
    d_me = torch.full((G,), -7, dtype=torch.int32, device=dev)
    d_ti = torch.full((G,), -7, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(2, dtype=torch.int32, device=dev)
    _C.compact_mask(grid.to(dev).view(torch.uint8), d_gi, d_me, d_cnt, pgi.to(dev), d_ti)
    E = me.numel()
    assert d_cnt.tolist() == [E, G - E]
    assert torch.equal(d_gi.cpu(), gi) and torch.equal(d_me[:E].cpu(), me) and torch.equal(d_ti[:G - E].cpu(), ti)
    d_me, d_ti = d_me[:E], d_ti[:G - E]

    d_plane = _fmt(plane.to(dev), nhwc)
    # ---- gather -------------------------------------------------------------------------------
    want = O.split(plane, me, BS)
    tiles = _fmt(torch.full((E, C, BS, BS), 7.0, dtype=dtype, device=dev), nhwc)
    _C.gather(tiles, d_plane, d_me, E)
    assert _same(tiles, want), "gather"
    # ---- gather with halo from the plane ------------------------------------------------------
    want_h = O.plane_halo(plane, me, BS, pad)
    padded = _fmt(torch.full((E, C, BS + 2 * pad, BS + 2 * pad), 7.0, dtype=dtype, device=dev), nhwc)
    _C.gather_halo(padded, d_plane, d_me, E, BS, pad)
    assert _same(padded, want_h), "gather_halo"
    # ---- scatter in place ---------------------------------------------------------------------
    new_tiles = _rand((E, C, BS, BS), dtype, g)
    want_s = prevp.clone()
    O.combine_(new_tiles, want_s, me)
    d_prev = _fmt(prevp.to(dev), nhwc)
    _C.scatter(_fmt(new_tiles.to(dev), nhwc), d_prev, d_me, E)
    assert _same(d_prev, want_s), "scatter"
    # ---- copy_blocks (non in place combine) ---------------------------------------------------
    d_prev2 = _fmt(prevp.to(dev), nhwc)
    d_out = _fmt(torch.full((N, C, H, W), 7.0, dtype=dtype, device=dev), nhwc)
    _C.copy_blocks(d_out, d_prev2, _fmt(new_tiles.to(dev), nhwc), d_gi)
    assert _same(d_out, want_s), "copy_blocks"
    assert _same(d_prev2, prevp), "copy_blocks must not touch prev"
    # ---- ring transfer (reference tile protocol) ----------------------------------------------
    Ep = pme.numel()
    prev_exec = _rand((Ep, C, BS, BS), dtype, g)
    prev_tr = _rand((G - Ep, C, BS, BS), dtype, g)
    base = _rand((G - E, C, BS, BS), dtype, g)
    want_t = base.clone()
    O.transfer(want_t, prev_exec, prev_tr, ti, G, pad)
    d_t = _fmt(base.to(dev), nhwc)
    _C.transfer(d_t, _fmt(prev_exec.to(dev), nhwc), _fmt(prev_tr.to(dev), nhwc), d_ti, G, pad)
    assert _same(d_t, want_t), "transfer"
    # ---- gather with halo, tile protocol ------------------------------------------------------
    want_r = O.repad(new_tiles, want_t, gi, me, pad)
    d_r = _fmt(torch.full((E, C, BS + 2 * pad, BS + 2 * pad), 7.0, dtype=dtype, device=dev), nhwc)
    _C.gather_halo_tiles(d_r, _fmt(new_tiles.to(dev), nhwc), d_t, d_gi, d_me, E, pad)
    assert _same(d_r, want_r), "gather_halo_tiles"
    torch.cuda.synchronize()


# (C, BS, pad) of every padded op of SwiftNet-RN18 at block 128 on a small 3x4 grid
SWIFTNET_LAYERS = [(3, 128, 3), (64, 64, 1), (64, 32, 1), (128, 16, 1), (256, 8, 1), (512, 4, 1), (128, 8, 1),
                   (128, 16, 1), (128, 32, 1), (19, 32, 1)]


@pytest.mark.parametrize("tma", [True, False])
@pytest.mark.parametrize("nhwc", [True, False])
@pytest.mark.parametrize("C,BS,pad", SWIFTNET_LAYERS)
def test_swiftnet_layer_shapes_fp16(C, BS, pad, nhwc, tma):
    from blockcopy import _C

    _C.set_tma_enabled(tma)
    try:
        _run_all(1, C, 3, 4, BS, pad, torch.float16, nhwc, 0.35, seed=C + BS)
    finally:
        _C.set_tma_enabled(True)


@pytest.mark.parametrize("nhwc", [True, False])
@pytest.mark.parametrize("N,C,GH,GW,BS,pad,frac", [
    (1, 8, 2, 2, 4, 1, 0.0),     # E = 0
    (1, 8, 2, 2, 4, 1, 1.0),     # E = all
    (2, 16, 3, 5, 8, 2, 0.5),    # N = 2, p = 2 (dilated Pedestron convs)
    (3, 8, 2, 3, 16, 3, 0.4),    # N = 3, p = 3
    (1, 24, 4, 4, 2, 1, 0.5),    # BS = 2
    (1, 32, 5, 3, 1, 1, 0.5),    # BS = 1 (deepest level of small inputs)
    (1, 5, 3, 3, 6, 1, 0.5),     # odd channel count, BS not a power of two
    (2, 7, 2, 2, 10, 2, 0.6),
])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_edge_cases(N, C, GH, GW, BS, pad, frac, dtype, nhwc):
    _run_all(N, C, GH, GW, BS, pad, dtype, nhwc, frac, seed=BS * 7 + C)


@pytest.mark.parametrize("tma", [True, False])
@pytest.mark.parametrize("nhwc", [False, True])
def test_baseline_config2(nhwc, tma):
    """BASELINE.json configs[1]: 1x128x256x512 fp16, 128-px image blocks (= 32 px at this level),
    grid 8x16, E = 38 of 128 (first 38 of randperm(128, seed 0)); bit-exact vs the oracle."""
    from blockcopy import _C

    dev = "cuda"
    X = torch.randn(1, 128, 256, 512, generator=torch.Generator().manual_seed(0)).half()
    P = torch.randn(1, 128, 256, 512, generator=torch.Generator().manual_seed(1)).half()
    cells = torch.randperm(128, generator=torch.Generator().manual_seed(0))[:38]
    grid = torch.zeros(128, dtype=torch.bool)
    grid[cells] = True
    grid = grid.view(1, 1, 8, 16)
    gi, me = O.grid_mappings(grid)
    E, BS = 38, 32
    _C.set_tma_enabled(tma)
    try:
        d_me, d_gi = me.to(dev), gi.to(dev)
        dX = _fmt(X.to(dev), nhwc)
        tiles = _fmt(torch.empty(E, 128, BS, BS, dtype=torch.float16, device=dev), nhwc)
        _C.gather(tiles, dX, d_me, E)
        want = O.split(X, me, BS)
        assert _same(tiles, want)
        padded = _fmt(torch.empty(E, 128, BS + 2, BS + 2, dtype=torch.float16, device=dev), nhwc)
        _C.gather_halo(padded, dX, d_me, E, BS, 1)
        assert _same(padded, O.plane_halo(X, me, BS, 1))
        dP = _fmt(P.to(dev), nhwc)
        out = torch.empty_like(dP)
        _C.copy_blocks(out, dP, tiles, d_gi)
        wantc = P.clone()
        O.combine_(want, wantc, me)
        assert _same(out, wantc)
        _C.scatter(tiles, dP, d_me, E)
        assert _same(dP, wantc)
    finally:
        _C.set_tma_enabled(True)


def test_full_size_roundtrip_properties():
    """Size-independent properties at config-5 size (2048x4096 image => 512x1024 plane, 134 MB):
    scatter(gather(X)) leaves X unchanged; gather(scatter(T)) == T; copy_blocks == clone+scatter;
    the halo gather's interior equals the plain gather and its frame border is zero."""
    from blockcopy import _C

    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    N, C, H, W, BS = 1, 128, 512, 1024, 32
    X = torch.randn(N, C, H, W, generator=g, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    grid = torch.rand(N, 1, H // BS, W // BS, generator=g, device=dev) < 0.3
    gi = torch.empty(grid.shape, dtype=torch.int32, device=dev)
    me = torch.empty(grid.numel(), dtype=torch.int32, device=dev)
    cnt = torch.empty(2, dtype=torch.int32, device=dev)
    _C.compact_mask(grid.view(torch.uint8), gi, me, cnt)
    E = int(cnt[0])
    assert E == int(grid.sum())
    me = me[:E]
    tiles = torch.empty(E, C, BS, BS, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    _C.gather(tiles, X, me, E)
    X2 = X.clone()
    _C.scatter(tiles, X2, me, E)
    assert torch.equal(X2, X)
    T = torch.randn(tiles.shape, generator=g, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    _C.scatter(T, X2, me, E)
    back = torch.empty_like(T)
    _C.gather(back, X2, me, E)
    assert torch.equal(back, T)
    out = torch.empty_like(X)
    _C.copy_blocks(out, X, T, gi)
    assert torch.equal(out, X2)
    padded = torch.empty(E, C, BS + 2, BS + 2, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    _C.gather_halo(padded, X2, me, E, BS, 1)
    assert torch.equal(padded[:, :, 1:-1, 1:-1], T)
    top = (me.long() % ((H // BS) * (W // BS))) < (W // BS)
    assert float(padded[top][:, :, 0, :].abs().sum()) == 0.0
    # checksum of checksums: halo ring of every tile == the plane's neighbouring strip
    ref = torch.nn.functional.pad(X2, (1, 1, 1, 1))
    cell = int(me[E // 2])
    gh, gw = divmod(cell, W // BS)
    assert torch.equal(padded[E // 2], ref[0, :, gh * BS:gh * BS + BS + 2, gw * BS:gw * BS + BS + 2])


def test_reference_named_functions_bind_to_the_library():
    """utils/block_funcs.py / blockpad.py keep the reference's call signatures."""
    from blockcopy.utils.block_funcs import CombineFunction, SplitFunction, TransferFunction
    from blockcopy.utils.blockpad import pad

    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    img = torch.randn(1, 6, 16, 24, generator=g).half()
    grid0 = torch.ones(1, 1, 2, 3, dtype=torch.bool)
    grid1 = torch.tensor([[1, 0, 1], [0, 1, 0]], dtype=torch.bool).view(1, 1, 2, 3)
    gi0, me0 = O.grid_mappings(grid0)
    gi1, me1 = O.grid_mappings(grid1)
    blocks0 = SplitFunction.apply(torch.empty(6, 6, 8, 8, dtype=torch.float16, device=dev), img.to(dev), me0.to(dev), gi0.to(dev))
    assert _same(blocks0, O.split(img, me0, 8))
    ti = O.transfer_idx(grid1, gi0)
    tr = TransferFunction.apply(torch.zeros(3, 6, 8, 8, dtype=torch.float16, device=dev), blocks0,
                                torch.empty(0, 6, 8, 8, dtype=torch.float16, device=dev), gi0.to(dev), ti.to(dev), 1)
    want_tr = torch.zeros(3, 6, 8, 8, dtype=torch.float16)
    O.transfer(want_tr, O.split(img, me0, 8), torch.empty(0, 6, 8, 8, dtype=torch.float16), ti, 6, 1)
    assert _same(tr, want_tr)
    new = torch.randn(3, 6, 8, 8, generator=g).half()
    padded = pad(new.to(dev), tr, gi1.to(dev), me1.to(dev), 1)
    assert _same(padded, O.repad(new, want_tr, gi1, me1, 1))
    out = CombineFunction.apply(new.to(dev), img.to(dev).clone(), gi1.to(dev), me1.to(dev))
    want = img.clone()
    O.combine_(new, want, me1)
    assert _same(out, want)
    with pytest.raises(AttributeError):  # plane not divisible by the block size
        from blockcopy import _C
        _C.gather(torch.empty(1, 6, 5, 5, dtype=torch.float16, device=dev), img.to(dev), me0.to(dev), 1)
