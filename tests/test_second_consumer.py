"""Second consumer of the API: the op set of Pedestron's CSPBlockCopy (dilated conv with padding 2,
ConvTranspose2d per block, channel L2 norm, cat, GroupNorm over all executed blocks, to_tensor in the middle
of the head, detector-side inlined state machine) -- tests/csp_standin.py -- against the fixture the
UNMODIFIED reference package produced for the same module (oracle/make_golden_cpu.py)."""
import os

import pytest
import torch

from cpu_backend import cpu_backend


def _run(device, dtype, golden_dir, wide=False, profile_last=False):
    import blockcopy
    from blockcopy.core.argparser import default_settings
    from consumers.clips import PolicyReplay, deterministic_init_, synthetic_clip
    from csp_standin import StandinDetector

    fix = torch.load(os.path.join(golden_dir, "csp_standin_wide_cpu.pt" if wide else "csp_standin_cpu.pt"))
    det = deterministic_init_(StandinDetector(default_settings(block_policy="all", block_size=fix["BS"]), wide=wide).eval(),
                              seed=fix["init_seed"])
    det = det.to(device=device, dtype=dtype)
    det.policy = PolicyReplay(fix["BS"], list(fix["grids"].bool()))
    clip = synthetic_clip(len(fix["grids"]), fix["H"], fix["W"], seed=fix["clip_seed"], dtype=dtype, device=device)
    with torch.no_grad():
        outs = [det.simple_test(f) for f in clip[:-1]]
        if profile_last:
            from torch.profiler import ProfilerActivity, profile

            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                outs.append(det.simple_test(clip[-1]))
                torch.cuda.synchronize()
            names = [e.key for e in prof.key_averages() if "cuda" in str(getattr(e, "device_type", "")).lower()]
            return fix, outs, names
        outs.append(det.simple_test(clip[-1]))
    return fix, outs


@pytest.mark.parametrize("wide", [False, True])
def test_csp_standin_cpu_matches_reference(golden_dir, wide):
    with cpu_backend():
        fix, outs = _run("cpu", torch.float32, golden_dir, wide)
    for t, o in enumerate(outs):
        assert type(o) is torch.Tensor
        assert torch.allclose(o, fix["outs"][t], atol=2e-4, rtol=1e-4), (t, (o - fix["outs"][t]).abs().max())
    assert outs[3] is outs[2]  # empty mask: previous output object


@pytest.mark.gpu
@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.float16, 4e-2)])
def test_csp_standin_gpu_matches_reference(golden_dir, dtype, tol, wide):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fix, outs = _run("cuda", dtype, golden_dir, wide)
    for t, o in enumerate(outs):
        ref = fix["outs"][t]
        err = (o.float().cpu() - ref).abs().max().item()
        assert err <= tol * float(ref.abs().max()), (t, err)


@pytest.mark.gpu
def test_csp_standin_wide_block_ops_run_on_native_kernels(golden_dir):
    """At CSP's channel structure (multiples of 64, 7x7 stem) every BLOCK op of a steady fp16 frame -- stem, strided /
    dilated 3x3 convs, both ConvTranspose2d, GroupNorm -- is one of this library's kernels: the only library
    convolution left is the DENSE 3x3 `cls` conv the head runs after `blockcopy.to_tensor` (outside the block path,
    csp_head.py:137-152), and no group-norm / transposed-conv kernel of ATen / cuDNN runs at all."""
    fix, outs, names = _run("cuda", torch.float16, golden_dir, wide=True, profile_last=True)
    ref = fix["outs"][-1]
    assert (outs[-1].float().cpu() - ref).abs().max().item() <= 4e-2 * float(ref.abs().max())
    ours = [n for n in names if "bc::" in n]
    for k in ("conv_stem", "conv_igemm", "gn_stats", "depth_to_space", "ew_fused"):
        assert any(k in n for n in ours), (k, ours)
    low = [n.lower() for n in names if "bc::" not in n]
    assert not [n for n in low if "group_norm" in n or "groupnorm" in n or "rowwisemoments" in n], low
    assert not [n for n in low if "dgrad" in n or "transpose" in n and "conv" in n], low
    lib_convs = [n for n in low if any(s in n for s in ("cudnn", "cutlass", "implicit", "gemm", "conv"))]
    assert len(lib_convs) <= 2, lib_convs  # the dense cls conv (+ its layout helper)
