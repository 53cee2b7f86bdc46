#!/usr/bin/env python
"""Per-CTA timelines of bc_conv_igemm / bc_conv_stem through bc_debug_trace (include/blockcopy_b200.h).
Prints, per layer shape of SwiftNet-RN18 at 1024x2048 / E=40: mean SM clocks spent per phase of a CTA,
and the wall-clock span of the launch from %globaltimer.   usage: python tools/cta_timeline.py [E]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))
import torch  # noqa: E402
from blockcopy import _C  # noqa: E402

dev = torch.device("cuda", 0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 40
PH = ["setup", "first operands", "k-loop issue", "drain->acc", "epilogue", "exit sync"]


def report(name, trace, flops=None):
    t = trace.cpu().view(-1, 16)
    t = t[t[:, 0] != 0]
    n = t.shape[0]
    c = t[:, :7].double()
    d = [(c[:, 1] - c[:, 0]), (c[:, 2] - c[:, 1]), (c[:, 3] - c[:, 2]), (c[:, 4] - c[:, 3]), (c[:, 5] - c[:, 4]),
         (c[:, 6] - c[:, 5])]
    life = c[:, 6] - c[:, 0]
    wall0, wall1 = t[:, 8].min().item(), t[:, 10].max().item()
    starts = (t[:, 8] - wall0).double() / 1e3
    sms = len(set(t[:, 9].tolist()))
    print(f"\n## {name}: {n} CTAs on {sms} SMs, launch span {1e-3 * (wall1 - wall0):.1f} us"
          + (f" ({flops / (wall1 - wall0) * 1e-3:.0f} TFLOP/s)" if flops else ""))
    print("   phase clocks (mean / p90): " + "; ".join(
        f"{p} {x.mean():.0f}/{x.quantile(0.9):.0f}" for p, x in zip(PH, d)))
    if (t[:, 11] != 0).any():
        m = [(t[:, k].double() - c[:, 1]).mean() for k in (11, 13, 14, 12)]
        print(f"   producer 0 after the PDL wait: tile decoded +{m[0]:.0f} clk, block coordinates (mapping loads) +{m[1]:.0f}, "
              f"tap decoded +{m[2]:.0f}, first activation load issued +{m[3]:.0f}, first operands landed +{d[1].mean():.0f}")
    if (t[:, 7] != 0).any() and (t[:, 4] != 0).any() and not (t[:, 13] != 0).any():
        d0 = t[:, 4].double() - c[:, 3]
        a = (t[:, 7] - t[:, 4]).double()
        b = c[:, 5] - t[:, 7].double()
        print(f"   epilogue of the last tile: last MMA issue -> accumulator complete {d0.mean():.0f} clk, TMEM -> staging rows "
              f"{a.mean():.0f} clk, staging -> global {b.mean():.0f} clk")
    if (t[:, 13] != 0).any():
        w = (t[:, 14] - t[:, 13]).double()
        r = c[:, 6] - t[:, 14].double()
        print(f"   split-K tail: thread 0 reaches the cluster barrier {(t[:, 13].double() - c[:, 3]).mean():.0f} clk after its last MMA issue, "
              f"waits {w.mean():.0f} (p90 {w.quantile(0.9):.0f}) clk in it, reduce + stores {r.mean():.0f} clk")
    print(f"   CTA lifetime clocks mean {life.mean():.0f} p90 {life.quantile(0.9):.0f}; "
          f"CTA start offsets us: p50 {starts.quantile(0.5):.1f} p90 {starts.quantile(0.9):.1f} max {starts.max():.1f}")


def main():
    trace = torch.zeros(16 * 8192, dtype=torch.int64, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    cells = torch.randperm(128, generator=torch.Generator().manual_seed(0))[:E].sort().values.to(torch.int32).to(dev)
    cl = torch.channels_last

    def run(fn, name, flops):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        trace.zero_()
        _C.lib().bc_debug_trace(trace.data_ptr())
        fn()
        torch.cuda.synchronize()
        _C.lib().bc_debug_trace(None)
        report(name, trace, flops)

    # stem: frame 1024x2048, block 128 -> BS_out 64
    BS = 128
    H, W = 8 * BS, 16 * BS
    tiles = torch.randn(E, 3, BS, BS, device=dev, dtype=torch.float16, generator=g)
    plane = _C.stem_plane(1, H // 2, W // 2, torch.float16, dev)
    _C.stem_pack(plane, tiles, cells, E)
    w7 = (torch.randn(64, 3, 7, 7, device=dev, dtype=torch.float16, generator=g) * 0.1)
    wp = _C.pack_stem_weight(w7)
    b = torch.zeros(64, device=dev, dtype=torch.float16)
    out = torch.empty(E, 64, BS // 2, BS // 2, device=dev, dtype=torch.float16).contiguous(memory_format=cl)
    nxt = torch.empty(1, 64, H // 2, W // 2, device=dev, dtype=torch.float16).contiguous(memory_format=cl)
    run(lambda: _C.conv_stem(out, plane, wp, b, cells, E, relu=True, plane_out=nxt), "stem 7x7/s2 -> 64ch, BS_out 64",
        2.0 * 147 * 64 * (BS // 2) ** 2 * E)

    for name, Cin, Cout, BS in (("conv3x3 c128 bs32 (#20)", 128, 128, 32), ("conv3x3 c64 bs32 (layer1)", 64, 64, 32),
                                ("conv3x3 c128 bs16 (layer2)", 128, 128, 16), ("conv3x3 c256 bs8 (layer3)", 256, 256, 8),
                                ("conv3x3 c512 bs4 (layer4)", 512, 512, 4)):
        H, W = 8 * BS, 16 * BS
        pl = torch.randn(1, Cin, H, W, device=dev, dtype=torch.float16, generator=g).contiguous(memory_format=cl)
        w = (torch.randn(Cout, Cin, 3, 3, device=dev, dtype=torch.float16, generator=g) * 0.05).contiguous(memory_format=cl)
        bias = torch.zeros(Cout, device=dev, dtype=torch.float16)
        o = torch.empty(E, Cout, BS, BS, device=dev, dtype=torch.float16).contiguous(memory_format=cl)
        nx = torch.empty(1, Cout, H, W, device=dev, dtype=torch.float16).contiguous(memory_format=cl)
        run(lambda: _C.conv_igemm(o, pl, w, bias, None, cells, E, BS, 1, 1, relu=True, plane_out=nx), name,
            2.0 * 9 * Cin * Cout * BS * BS * E)


if __name__ == "__main__":
    main()
