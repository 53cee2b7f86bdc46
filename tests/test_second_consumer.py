"""Second consumer of the API: the op set of Pedestron's CSPBlockCopy (dilated conv with padding 2,
ConvTranspose2d per block, channel L2 norm, cat, GroupNorm over all executed blocks, to_tensor in the middle
of the head, detector-side inlined state machine) -- tests/csp_standin.py -- against the fixture the
UNMODIFIED reference package produced for the same module (oracle/make_golden_cpu.py)."""
import os

import pytest
import torch

from cpu_backend import cpu_backend


def _run(device, dtype, golden_dir):
    import blockcopy
    from blockcopy.core.argparser import default_settings
    from consumers.clips import PolicyReplay, deterministic_init_, synthetic_clip
    from csp_standin import StandinDetector

    fix = torch.load(os.path.join(golden_dir, "csp_standin_cpu.pt"))
    det = deterministic_init_(StandinDetector(default_settings(block_policy="all", block_size=fix["BS"])).eval(),
                              seed=fix["init_seed"])
    det = det.to(device=device, dtype=dtype)
    det.policy = PolicyReplay(fix["BS"], list(fix["grids"].bool()))
    clip = synthetic_clip(len(fix["grids"]), fix["H"], fix["W"], seed=fix["clip_seed"], dtype=dtype, device=device)
    with torch.no_grad():
        outs = [det.simple_test(f) for f in clip]
    return fix, outs


def test_csp_standin_cpu_matches_reference(golden_dir):
    with cpu_backend():
        fix, outs = _run("cpu", torch.float32, golden_dir)
    for t, o in enumerate(outs):
        assert type(o) is torch.Tensor
        assert torch.allclose(o, fix["outs"][t], atol=2e-4, rtol=1e-4), (t, (o - fix["outs"][t]).abs().max())
    assert outs[3] is outs[2]  # empty mask: previous output object


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.float16, 4e-2)])
def test_csp_standin_gpu_matches_reference(golden_dir, dtype, tol):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fix, outs = _run("cuda", dtype, golden_dir)
    for t, o in enumerate(outs):
        ref = fix["outs"][t]
        err = (o.float().cpu() - ref).abs().max().item()
        assert err <= tol * float(ref.abs().max()), (t, err)
