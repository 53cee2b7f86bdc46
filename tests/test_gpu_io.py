"""bc_frame_from_u8 / bc_upsample_argmax against the oracle (bit-exact) and against the torch op sequences of
the reference's driver on the GPU (lib/ext_transforms.py:317-372, test_swiftnet.py:196-197)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import cpu_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("N,H,W", [(1, 64, 128), (2, 24, 40), (1, 7, 9), (1, 1024, 2048)])
def test_frame_from_u8_bit_exact(dtype, N, H, W):
    from consumers.frame_io import CITYSCAPES_MEAN as M, CITYSCAPES_STD as S, FrameNormalizer

    g = torch.Generator().manual_seed(H * W + N)
    u8 = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, generator=g)
    got = FrameNormalizer(dtype=dtype)(u8.cuda())
    assert got.shape == (N, 3, H, W) and got.dtype == dtype
    assert torch.equal(got.cpu(), cpu_oracle.frame_from_u8(u8, M, S, dtype))
    # and the torch sequence of the driver.  It runs on the CPU there (DataLoader workers): CPU `div(255)` is a
    # true division, while torch's CUDA kernel multiplies by the reciprocal (1 ulp apart for some bytes)
    t = u8.permute(0, 3, 1, 2).contiguous().float().div(255)
    mean = torch.as_tensor(M, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.as_tensor(S, dtype=torch.float32).view(1, 3, 1, 1)
    assert torch.equal(got.cpu(), t.sub_(mean).div_(std).to(dtype))


def test_frame_from_u8_unbatched_and_errors():
    from blockcopy import _C

    u8 = torch.randint(0, 256, (32, 48, 3), dtype=torch.uint8)
    got = _C.frame_from_u8(u8.cuda(), (0.5, 0.5, 0.5), (0.25, 0.5, 1.0), torch.float32)
    assert torch.equal(got.cpu(), cpu_oracle.frame_from_u8(u8[None], (0.5, 0.5, 0.5), (0.25, 0.5, 1.0), torch.float32))
    with pytest.raises(_C.BlockCopyNativeError):
        _C.frame_from_u8(u8.cuda(), (0, 0, 0), (1, 0, 1))  # std of zero
    with pytest.raises(AssertionError):
        _C.frame_from_u8(u8, (0, 0, 0), (1, 1, 1))  # CPU tensor: no fallback


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("label", [torch.uint8, torch.int64])
@pytest.mark.parametrize("N,K,h,w,s,cl", [(1, 19, 32, 64, 4, False), (2, 19, 16, 24, 4, True), (1, 5, 9, 7, 2, False),
                                          (1, 3, 6, 5, 1, False), (1, 19, 1, 8, 4, False), (1, 19, 8, 1, 4, True)])
def test_upsample_argmax_equals_oracle(dtype, label, N, K, h, w, s, cl):
    from blockcopy import _C

    g = torch.Generator().manual_seed(K * h + w)
    x = (2 * torch.randn(N, K, h, w, generator=g)).to(dtype)
    xd = x.cuda()
    if cl:
        xd = xd.contiguous(memory_format=torch.channels_last)
    got = _C.upsample_argmax(xd, s, label)
    want = cpu_oracle.upsample_argmax(x, s)
    assert got.dtype == label and got.shape == want.shape
    if dtype == torch.float16:
        assert torch.equal(got.cpu().long(), want)
    else:  # fp32: a fused multiply-add may move a blended value by one ulp
        diff = got.cpu().long() != want
        assert diff.float().mean() < 1e-3


def test_upsample_argmax_full_size_vs_torch_sequence():
    """BASELINE size: (1,19,256,512) fp16 logits -> 1024x2048 labels vs F.interpolate + max on the GPU."""
    from consumers.frame_io import predict_labels

    g = torch.Generator().manual_seed(11)
    x = (3 * torch.randn(1, 19, 256, 512, generator=g)).half().cuda()
    got = predict_labels(x, size=(1024, 2048), label_dtype=torch.int64)
    up = F.interpolate(x, size=(1024, 2048), mode="bilinear")
    vals, want = up.max(dim=1)
    # the chosen class always holds the maximum value; where the maximum is unique the index is the same
    assert torch.equal(up.gather(1, got[:, None])[:, 0], vals)
    unique = (up == vals[:, None]).sum(1) == 1
    assert torch.equal(got[unique], want[unique])
    assert unique.float().mean() > 0.99
    # uint8 labels, default scale
    assert torch.equal(predict_labels(x).long(), got)
    # ties -> lowest class
    assert int(predict_labels(torch.zeros(1, 19, 8, 8, device="cuda").half()).max()) == 0


def test_upsample_argmax_errors():
    from blockcopy import _C

    x = torch.randn(1, 19, 8, 8).half().cuda()
    with pytest.raises(_C.BlockCopyNativeError):
        _C.upsample_argmax(x, 3)
    with pytest.raises(_C.BlockCopyNativeError):
        _C.upsample_argmax(torch.randn(1, 300, 4, 4).half().cuda(), 4, torch.uint8)
    with pytest.raises(AssertionError):
        _C.upsample_argmax(x.cpu(), 4)


def test_kernels_equal_committed_golden(golden_dir):
    """The CUDA kernels against tests/golden/io_kat.pt (made by torchvision / torch CPU calls, not by the oracle)."""
    import os

    from blockcopy import _C

    fix = torch.load(os.path.join(golden_dir, "io_kat.pt"))
    for dtype, key in ((torch.float32, "frames_fp32"), (torch.float16, "frames_fp16")):
        got = _C.frame_from_u8(fix["u8"].cuda(), fix["mean"], fix["std"], dtype)
        assert torch.equal(got.cpu(), fix[key])
    for lab in (torch.uint8, torch.int64):
        assert torch.equal(_C.upsample_argmax(fix["logits16"].cuda(), 4, lab).cpu().long(), fix["labels16"])
    got32 = _C.upsample_argmax(fix["logits32"].cuda(), 4, torch.int64).cpu()
    diff = got32 != fix["labels32"]
    assert not diff.any() or float(fix["top2_gap32"][diff].max()) < 1e-6


@pytest.mark.parametrize("N,K,h,w,GH,GW,scale,frac", [(1, 19, 256, 512, 8, 16, 4, 0.3), (2, 19, 64, 128, 4, 8, 4, 0.5),
                                                       (1, 7, 32, 32, 4, 4, 2, 0.2), (1, 19, 64, 64, 2, 2, 4, 1.0),
                                                       (1, 19, 48, 96, 3, 6, 4, 0.0), (1, 3, 8, 8, 8, 8, 4, 0.4)])
def test_block_sparse_label_update_equals_dense_argmax(N, K, h, w, GH, GW, scale, frac):
    """bc_upsample_argmax_blocks: the previous frame's label map + this frame's logits (changed only inside executed
    cells) -> exactly the label map bc_upsample_argmax computes from scratch (1-px logit blocks included)."""
    from blockcopy import _C
    from consumers.frame_io import BlockLabelMap

    g = torch.Generator().manual_seed(h + GH)
    prev = (2 * torch.randn(N, K, h, w, generator=g)).half()
    grid = torch.rand(N, 1, GH, GW, generator=g) < frac
    BS = h // GH
    mask = grid.repeat_interleave(BS, 2).repeat_interleave(BS, 3)
    cur = torch.where(mask, (2 * torch.randn(N, K, h, w, generator=g)).half(), prev)
    lm = BlockLabelMap(scale=scale)
    first = lm.update(prev.cuda(), None)
    assert torch.equal(first, _C.upsample_argmax(prev.cuda(), scale))
    ptr = first.data_ptr()
    got = lm.update(cur.cuda(), grid.cuda())
    assert got.data_ptr() == ptr  # updated in place
    want = _C.upsample_argmax(cur.cuda(), scale)
    assert torch.equal(got, want), int((got != want).sum())
    if 0 < frac < 1:
        assert not torch.equal(want, _C.upsample_argmax(prev.cuda(), scale))  # the update was needed


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("N,GH,GW,BS,frac", [(1, 4, 6, 64, 0.4), (2, 3, 2, 128, 0.5), (1, 8, 16, 16, 0.3), (1, 2, 2, 32, 1.0)])
def test_blocks_from_u8_equals_oracle_normalise_then_split(dtype, N, GH, GW, BS, frac):
    """bc_blocks_from_u8 == the oracle's frame normalisation followed by the oracle's split (reference
    ext_transforms.py:317-372 then tensorwrapper.py:335-381), bit for bit; and == bc_frame_from_u8 + bc_gather."""
    from blockcopy import _C

    g = torch.Generator().manual_seed(BS + GH)
    H, W = GH * BS, GW * BS
    u8 = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, generator=g)
    grid = torch.rand(N, 1, GH, GW, generator=g) < frac
    grid[0, 0, 0, 0] = True
    _, me = cpu_oracle.grid_mappings(grid)
    E = me.numel()
    M, S = (0.287, 0.325, 0.284), (0.176, 0.181, 0.178)
    want = cpu_oracle.split(cpu_oracle.frame_from_u8(u8, M, S, dtype).contiguous(), me, BS)
    tiles = torch.full((E, 3, BS, BS), float("nan"), dtype=dtype, device="cuda")
    _C.blocks_from_u8(tiles, u8.cuda(), M, S, me.cuda(), E)
    assert torch.equal(tiles.cpu(), want)
    dense = _C.frame_from_u8(u8.cuda(), M, S, dtype)
    two = torch.empty_like(tiles)
    _C.gather(two, dense, me.cuda(), E)
    assert torch.equal(two, tiles)


def test_model_on_u8_frames_equals_model_on_normalised_frames():
    """BlockCopyModel(U8Frame) in CUDA-graph mode: the steady frames' first gather reads the uint8 frame
    (bc_blocks_from_u8, no full-frame normalisation) and every output equals the one for the normalised tensor."""
    import blockcopy
    from blockcopy import _C
    from blockcopy.core.argparser import default_settings
    from blockcopy.core.frame import U8Frame
    from consumers.clips import PolicyFixedFraction
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    H, W, BS, T = 256, 512, 64, 7
    g = torch.Generator().manual_seed(3)
    base = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, generator=g)
    clip = [base]
    for _ in range(T - 1):
        nxt = clip[-1].clone()
        m = torch.rand(H, W, 1, generator=g) < 0.2
        nxt = torch.where(m, torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, generator=g), nxt)
        clip.append(nxt)
    clip = [f.cuda() for f in clip]
    outs, calls = {}, {"fused": 0, "dense": 0}
    orig_b, orig_f = _C.blocks_from_u8, _C.frame_from_u8
    _C.blocks_from_u8 = lambda *a, **k: (calls.__setitem__("fused", calls["fused"] + 1), orig_b(*a, **k))[1]
    _C.frame_from_u8 = lambda *a, **k: (calls.__setitem__("dense", calls["dense"] + 1), orig_f(*a, **k))[1]
    try:
        for mode in ("u8", "dense"):
            settings = default_settings(block_policy="all", block_size=BS)
            settings["block_cuda_graphs"] = True
            model = blockcopy.BlockCopyModel(build_swiftnet_rn18(seed=1), settings).eval().cuda().half()
            model.policy = PolicyFixedFraction(BS, fraction=0.3, quantize=2, seed=0)
            res = []
            with torch.no_grad():
                for rep in range(3):  # eager, capture, replay
                    model.reset_temporal()
                    if mode == "u8" and rep == 2:
                        calls["fused"] = calls["dense"] = 0
                    for f in clip:
                        x = U8Frame(f) if mode == "u8" else orig_f(f, U8Frame(f).mean, U8Frame(f).std, torch.float16)
                        res.append(model(x).clone())
                    if mode == "u8" and rep == 2:
                        # replayed clip: only the all-blocks first frame may need the whole normalised frame
                        assert calls["fused"] >= T - 2 and calls["dense"] <= 1, calls
            outs[mode] = res
    finally:
        _C.blocks_from_u8, _C.frame_from_u8 = orig_b, orig_f
    for a, b in zip(outs["u8"], outs["dense"]):
        assert torch.equal(a, b)
