"""The reference's OWN SwiftNet files (semantic_segmentation/lib/models/swiftnet/*.py + lib/utils/bn_fusion.py,
staged UNMODIFIED under baseline/_ref by __graft_entry__.build()) running on THIS blockcopy package -- the
north star's "SwiftNet calls it unchanged".  They import `blockcopy` (decorator, timings) and are wrapped by this
package's BlockCopyModel exactly as the reference driver does (test_swiftnet.py:107-123).

* CPU leg (`-m "not gpu"`): host logic over the oracle-backed kernels, fp32: reference module tree == this repo's
  consumers/swiftnet_rn18.py with the same state_dict, bit for bit, same number of native launches per frame.
* GPU leg (`-m gpu`): fp16, eager and CUDA-graph mode at 1024x2048 / 128-px blocks / E = 40 (the benchmarked
  configuration): same bits as consumers/swiftnet_rn18.py, the fused SPP / head / stem paths are taken (45 native
  launches per steady frame) and NO cuDNN / ATen convolution kernel runs in a steady frame.
"""
import contextlib
import io
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SS = os.path.join(ROOT, "baseline", "_ref", "semantic_segmentation")


def _reference_swiftnet():
    """The reference's SwiftNet(resnet18) module tree, eval mode, random init (its constructors print a lot)."""
    if not os.path.isdir(os.path.join(REF_SS, "lib", "models", "swiftnet")):
        pytest.skip("baseline/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
    import blockcopy  # noqa: F401  -- THIS package: the reference model files import it

    assert "blockcopy-video-processing-pytorch_b200" in blockcopy.__file__
    if REF_SS not in sys.path:
        sys.path.insert(0, REF_SS)
    with contextlib.redirect_stdout(io.StringIO()):
        from lib.models.swiftnet.backbones.resnet import resnet18
        from lib.models.swiftnet.swiftnet import SwiftNet
        from lib.utils import bn_fusion

        net = SwiftNet(resnet18(pretrained=False), num_classes=19, num_features=128, use_spp=True).eval()
    assert os.path.realpath(sys.modules["lib.models.swiftnet.swiftnet"].__file__).startswith(os.path.realpath(REF_SS))
    return net, bn_fusion


def _pair(settings, device, half, init_seed=0, gain=0.8):
    """(reference module tree, this repo's consumer) behind BlockCopyModel with identical weights; both are
    BN-fused by the reference's own bn_fusion.fuse_bn_recursively AFTER wrapping, like its driver."""
    import blockcopy
    from consumers.clips import deterministic_init_
    from consumers.swiftnet_rn18 import SwiftNetRN18

    ref_net, bn_fusion = _reference_swiftnet()
    deterministic_init_(ref_net, seed=init_seed, gain=gain)
    ours_net = SwiftNetRN18().eval()
    missing = ours_net.load_state_dict(ref_net.state_dict(), strict=True)  # identical parameter names
    assert not missing.missing_keys and not missing.unexpected_keys
    m_ref = blockcopy.BlockCopyModel(ref_net, dict(settings)).eval()
    with contextlib.redirect_stdout(io.StringIO()):
        m_ref = bn_fusion.fuse_bn_recursively(m_ref)
    m_ours = blockcopy.BlockCopyModel(ours_net, dict(settings)).eval()
    with contextlib.redirect_stdout(io.StringIO()):
        m_ours = bn_fusion.fuse_bn_recursively(m_ours)  # the same folding arithmetic => the same weight bits
    m_ref, m_ours = m_ref.to(device), m_ours.to(device)
    if half:
        m_ref, m_ours = m_ref.half(), m_ours.half()
    return m_ref, m_ours


def test_reference_swiftnet_files_run_on_this_package_cpu():
    from blockcopy import _C
    from blockcopy.core.argparser import default_settings
    from consumers.clips import PolicyFixedFraction, synthetic_clip
    from cpu_backend import cpu_backend

    BS, H, W, T = 32, 64, 128, 4
    m_ref, m_ours = _pair(default_settings(block_policy="all", block_size=BS), "cpu", half=False)
    clip = synthetic_clip(T, H, W, seed=2, dtype=torch.float32)
    outs, counts = {}, {}
    with cpu_backend(), torch.no_grad():
        for name, m in (("ref", m_ref), ("ours", m_ours)):
            m.policy = PolicyFixedFraction(BS, fraction=0.4, quantize=2, seed=3)
            m.reset_temporal()
            outs[name] = []
            for f in clip:
                outs[name].append((m(f).clone(), m.policy_meta["frame_state"].clone()))
            counts[name] = len(m.block_temporal_features._planes)
    assert counts["ref"] == counts["ours"] == 21  # padded ops of SwiftNet-RN18, found by call order
    for t, ((a, fa), (b, fb)) in enumerate(zip(outs["ref"], outs["ours"])):
        assert tuple(a.shape) == (1, 19, H // 4, W // 4)
        assert torch.equal(a, b), (t, float((a - b).abs().max()))
        assert torch.equal(fa, fb)
    assert _C is not None


def _steady_frames(model, clip, n_frames):
    """Run `n_frames` of the clip (frame 0 executes every block), return the clones of the outputs."""
    outs = []
    with torch.no_grad():
        model.reset_temporal()
        if hasattr(model.policy, "reseed"):
            model.policy.reseed(0)
        for f in clip[:n_frames]:
            outs.append(model(f).clone())
    torch.cuda.synchronize()
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("graphs", [False, True])
def test_reference_swiftnet_files_run_on_this_package_gpu(graphs):
    """Benchmarked configuration: 1024x2048, 128-px blocks, frame 0 all blocks then 40 of 128."""
    from blockcopy import _C
    from blockcopy.core.argparser import default_settings
    from consumers.clips import PolicyFixedFraction, synthetic_clip

    H, W, BS, T = 1024, 2048, 128, 6
    settings = default_settings(block_policy="all", block_size=BS)
    settings["block_cuda_graphs"] = graphs
    m_ref, m_ours = _pair(settings, "cuda", half=True)
    clip = synthetic_clip(T, H, W, seed=5, dtype=torch.float16, device="cuda")
    res = {}
    for name, m in (("ref", m_ref), ("ours", m_ours)):
        m.policy = PolicyFixedFraction(BS, fraction=0.3, quantize=8, seed=0)
        for _ in range(3 if graphs else 1):  # graph mode: eager, capture, replay
            outs = _steady_frames(m, clip, T)
        n0 = _C.launch_count()
        with torch.no_grad():
            m(clip[T - 1])  # one more steady frame (E = 40), counted
        torch.cuda.synchronize()
        res[name] = (outs, _C.launch_count() - n0)
    for t, (a, b) in enumerate(zip(res["ref"][0], res["ours"][0])):
        assert torch.isfinite(a).all()
        assert torch.equal(a, b), (t, float((a.float() - b.float()).abs().max()))
    # fused stem / SPP / head and the lazy epilogue fusion are taken for the reference's module tree too
    assert res["ref"][1] == res["ours"][1], res
    assert res["ref"][1] <= 45, res


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,BS", [(1024, 2048, 128), (256, 512, 64)])
def test_reference_swiftnet_steady_frame_has_no_library_conv_kernels(H, W, BS):
    """Every kernel of a steady block-sparse frame of the reference's SwiftNet is one of this library's `bc::`
    kernels (plus torch's tiny fill / copy helpers): no cuDNN / cuBLAS / CUTLASS / ATen convolution, pooling,
    batch-norm or interpolation kernel."""
    from torch.profiler import ProfilerActivity, profile

    from blockcopy.core.argparser import default_settings
    from consumers.clips import PolicyFixedFraction, synthetic_clip

    # 256x512 with 64-px blocks (the reference's --block-size 64, also what smoke() runs) puts layer4 at 2-px blocks
    m_ref, _ = _pair(default_settings(block_policy="all", block_size=BS), "cuda", half=True)
    m_ref.policy = PolicyFixedFraction(BS, fraction=0.3, quantize=8 if BS == 128 else 2, seed=0)
    clip = synthetic_clip(4, H, W, seed=5, dtype=torch.float16, device="cuda")
    _steady_frames(m_ref, clip, 3)
    with profile(activities=[ProfilerActivity.CUDA]) as prof, torch.no_grad():
        m_ref(clip[3])
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if getattr(e, "device_type", None) is not None
             and "cuda" in str(e.device_type).lower()]
    assert names, "profiler saw no CUDA kernels"
    ours = [n for n in names if "bc::" in n]
    banned = [n for n in names if any(s in n.lower() for s in ("cudnn", "cutlass", "cublas", "gemm", "implicit",
                                                               "conv", "batch_norm", "upsample", "pool"))
              and "bc::" not in n]
    assert not banned, banned
    assert len(ours) >= 8, names
