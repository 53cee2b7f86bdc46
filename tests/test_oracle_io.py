"""CPU: the oracle's driver-side steps (oracle/cpu_oracle.py: frame_from_u8, upsample_argmax) pinned against the
torch / torchvision calls the reference's driver makes (lib/ext_transforms.py:317-372, test_swiftnet.py:196-197)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import cpu_oracle

MEAN = (73.1584 / 255, 82.9090 / 255, 72.3924 / 255)   # lib/datasets/cityscapes_vid.py:29-30
STD = (44.9149 / 255, 46.1529 / 255, 45.3192 / 255)


def _driver_normalize(u8_hwc: np.ndarray) -> torch.Tensor:
    """ExtToTensor + ExtNormalize, spelled with the tensor ops torchvision's to_tensor / normalize run."""
    t = torch.from_numpy(u8_hwc).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    mean = torch.as_tensor(MEAN, dtype=torch.float32).view(-1, 1, 1)
    std = torch.as_tensor(STD, dtype=torch.float32).view(-1, 1, 1)
    return t.sub_(mean).div_(std)


def test_frame_from_u8_equals_torchvision_sequence():
    rng = np.random.default_rng(0)
    u8 = rng.integers(0, 256, size=(2, 24, 40, 3), dtype=np.uint8)
    u8[0, 0, :, 0] = np.arange(40) * 6  # spread of values incl. 0 and > 234
    u8[0, 1, :16] = 255
    want32 = torch.stack([_driver_normalize(f) for f in u8])
    try:
        import torchvision.transforms.functional as TF
        from PIL import Image

        tv = torch.stack([TF.normalize(TF.to_tensor(Image.fromarray(f)), MEAN, STD) for f in u8])
        assert torch.equal(tv, want32), "the spelled-out sequence is not what torchvision computes"
    except ImportError:
        pass
    for dtype in (torch.float32, torch.float16):
        got = cpu_oracle.frame_from_u8(torch.from_numpy(u8), MEAN, STD, dtype)
        assert got.dtype == dtype and torch.equal(got, want32.to(dtype))


def test_all_256_values_per_channel():
    u8 = np.zeros((1, 16, 16, 3), dtype=np.uint8)
    u8[0, :, :, :] = np.arange(256, dtype=np.uint8).reshape(16, 16, 1)
    got = cpu_oracle.frame_from_u8(torch.from_numpy(u8), MEAN, STD, torch.float32)
    assert torch.equal(got[0], _driver_normalize(u8[0]))


def test_upsample_argmax_equals_interpolate_max_fp32():
    g = torch.Generator().manual_seed(3)
    for (N, K, h, w, s) in [(1, 19, 8, 12, 4), (2, 5, 7, 5, 2), (1, 3, 4, 4, 1), (1, 19, 1, 9, 4)]:
        x = torch.randn(N, K, h, w, generator=g)
        up = F.interpolate(x, size=(h * s, w * s), mode="bilinear")
        want = up.max(dim=1)[1]
        got = cpu_oracle.upsample_argmax(x, s)
        # fp32: FMA contraction inside ATen may move a value by one ulp; accept only pixels whose two best are that close
        diff = got != want
        if diff.any():
            top2 = up.topk(2, dim=1).values
            assert ((top2[:, 0] - top2[:, 1])[diff] <= 4e-7 * top2[:, 0].abs()[diff] + 1e-7).all()
        assert diff.float().mean() < 1e-3


def test_upsample_argmax_fp16_rounding_and_ties():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 19, 16, 24, generator=g).half()
    up = F.interpolate(x.float(), size=(64, 96), mode="bilinear").half()  # value the fp16 tensor would hold
    want = up.float().argmax(dim=1)  # first index among equal maxima
    vals, _ = up.float().max(dim=1)
    got = cpu_oracle.upsample_argmax(x, 4)
    # equal wherever the maximum is unique; at ties the lowest class index
    assert torch.equal(up.float().gather(1, got[:, None])[:, 0], vals)
    ties = (up.float() == vals[:, None]).sum(1) > 1
    first = (up.float() == vals[:, None]).float().argmax(dim=1)
    assert torch.equal(got[ties], first[ties])
    assert torch.equal(got[~ties], want[~ties])
    # constant logits: everything ties -> class 0
    assert int(cpu_oracle.upsample_argmax(torch.zeros(1, 4, 3, 3).half(), 4).max()) == 0


def test_oracle_equals_committed_golden(golden_dir):
    """tests/golden/io_kat.pt: outputs of torchvision / torch CPU calls (oracle/make_golden_io.py)."""
    import os

    fix = torch.load(os.path.join(golden_dir, "io_kat.pt"))
    for dtype, key in ((torch.float32, "frames_fp32"), (torch.float16, "frames_fp16")):
        assert torch.equal(cpu_oracle.frame_from_u8(fix["u8"], fix["mean"], fix["std"], dtype), fix[key])
    assert torch.equal(cpu_oracle.upsample_argmax(fix["logits16"], 4), fix["labels16"])
    got32 = cpu_oracle.upsample_argmax(fix["logits32"], 4)
    diff = got32 != fix["labels32"]
    assert not diff.any() or float(fix["top2_gap32"][diff].max()) < 1e-6  # fp32: only exact near-ties may differ
