"""Per-block ConvTranspose2d and GroupNorm at Pedestron CSP sizes (1024x2048 frame, 128-px image blocks, E = 40):
this library's kernels against torch (cuDNN / ATen) on the same fp16 channels_last tile batch.  us per call."""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
from blockcopy import _C  # noqa: E402


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    E, res = 40, {}
    torch.backends.cudnn.benchmark = True
    for (Cin, Cout, BS, k, s, p) in [(512, 256, 16, 4, 2, 1), (1024, 256, 8, 4, 4, 0), (2048, 256, 8, 4, 4, 0), (64, 64, 32, 4, 2, 1)]:
        x = torch.randn(E, Cin, BS, BS, device="cuda").half().contiguous(memory_format=torch.channels_last)
        w = (torch.randn(Cin, Cout, k, k, device="cuda") * 0.02).half()
        b = torch.randn(Cout, device="cuda").half()
        wp, bp = _C.pack_deconv_weight(w, b, s)
        ph = torch.empty(E, s * s * Cout, BS, BS, dtype=torch.float16, device="cuda").contiguous(memory_format=torch.channels_last)
        out = torch.empty(E, Cout, s * BS, s * BS, dtype=torch.float16, device="cuda").contiguous(memory_format=torch.channels_last)

        def ours():
            _C.conv_igemm(ph, x, wp, bp, None, None, E, BS, 1, wp.shape[2] // 2)
            _C.depth_to_space(out, ph, s)

        t_ours = timed(ours)
        t_conv = timed(lambda: _C.conv_igemm(ph, x, wp, bp, None, None, E, BS, 1, wp.shape[2] // 2))
        wcl = w.contiguous(memory_format=torch.channels_last)
        t_torch = timed(lambda: F.conv_transpose2d(x, wcl, b, stride=s, padding=p))
        ref = F.conv_transpose2d(x.float(), w.float(), b.float(), stride=s, padding=p)
        err = float((out.float() - ref).abs().max() / ref.abs().max())
        res[f"deconv_{Cin}_{Cout}_bs{BS}_k{k}s{s}"] = dict(ours_us=round(t_ours, 1), conv_part_us=round(t_conv, 1),
                                                          torch_us=round(t_torch, 1), rel_err=err)
    for (C, BS, G) in [(256, 32, 32), (64, 32, 8)]:
        x = torch.randn(E, C, BS, BS, device="cuda").half().contiguous(memory_format=torch.channels_last)
        w, b = torch.rand(C, device="cuda").half(), torch.randn(C, device="cuda").half()
        wf, bf = w.float(), b.float()
        ws = torch.zeros(_C.GN_STATS_WORKSPACE, dtype=torch.uint8, device="cuda")
        st = torch.empty(2, C, device="cuda")
        out = torch.empty_like(x)

        def ours():
            _C.gn_stats(x, G, 1e-5, st[0], st[1], ws)
            _C.ew_fused(out, x, None, (st[0], st[1], wf, bf), True)

        def ref():  # the reference's fold (tensorwrapper.py:600-633) + ReLU
            y = x.permute(1, 0, 2, 3).reshape(1, C, E * BS, BS)
            y = F.group_norm(y, G, w, b, 1e-5)
            return F.relu_(y.reshape(C, E, BS, BS).permute(1, 0, 2, 3))

        res[f"groupnorm_relu_{C}_bs{BS}_g{G}"] = dict(ours_us=round(timed(ours), 1), torch_us=round(timed(ref), 1),
                                                      stats_us=round(timed(lambda: _C.gn_stats(x, G, 1e-5, st[0], st[1], ws)), 1),
                                                      bytes_mb=round(x.numel() * 2 * 3 / 1e6, 1))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
