#!/usr/bin/env python
"""BASELINE config 5: active-fraction sweep at 2048x4096 (and 1024x2048), frames/s of the BlockCopy path
(CUDA-graph mode, fixed-fraction seeded masks, 30-frame clips) next to the dense cuDNN fp16 forward of the same
SwiftNet-RN18 (channels_last, eager and as one CUDA graph).  Device-resident inputs, CUDA-event timing.
usage: python tools/fraction_sweep.py [--sizes 1024x2048,2048x4096] > profiles/..."""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
from consumers.clips import synthetic_clip  # noqa: E402
from consumers.swiftnet_rn18 import build_swiftnet_rn18  # noqa: E402

FRACTIONS = [0.05, 0.10, 0.20, 0.30, 0.50, 0.75, 1.00]


def timed(fn, iters):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn(iters)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters  # ms per frame


def dense_fps(H, W, dev):
    net = build_swiftnet_rn18(seed=0).eval().to(dev).half().to(memory_format=torch.channels_last)
    x = torch.randn(1, 3, H, W, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        for _ in range(5):
            net(x)
        eager = timed(lambda n: [net(x) for _ in range(n)], 20)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(x)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            net(x)
        graphed = timed(lambda n: [g.replay() for _ in range(n)], 20)
    torch.backends.cudnn.benchmark = False
    del net, g
    torch.cuda.empty_cache()
    return 1e3 / eager, 1e3 / graphed


def blockcopy_fps(H, W, frac, dev, clip_len=30):
    args = types.SimpleNamespace(policy="fixed", fraction=frac, no_graphs=False)
    model = bench.build_model(args, dev)
    clip = synthetic_clip(clip_len, H, W, seed=0, dtype=torch.float16, device=dev)
    for warm in range(2):  # two clips: eager first occurrence of each block count, capture on the second
        bench.run_frames([model], [clip], warm * clip_len, clip_len, clip_len)
    ms = timed(lambda n: bench.run_frames([model], [clip], 2 * clip_len, n, clip_len), 2 * clip_len)
    meta = model.policy_meta
    active, total = int(meta["num_exec"]), int(meta["num_total"])
    del model, clip
    torch.cuda.empty_cache()
    return 1e3 / ms, active, total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024x2048,2048x4096")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    for size in a.sizes.split(","):
        H, W = (int(v) for v in size.split("x"))
        e, g = dense_fps(H, W, dev)
        print(f"\\n## {H}x{W}: dense cuDNN fp16 forward (channels_last) {e:.0f} frames/s eager, {g:.0f} frames/s as one CUDA graph\\n")
        print("| target fraction | executed / total blocks (steady frame) | BlockCopy frames/s (30-frame clips incl. the full first frame) | vs dense (graph) |")
        print("|---:|---:|---:|---:|")
        for f in FRACTIONS:
            fps, act, tot = blockcopy_fps(H, W, f, dev)
            print(f"| {100 * f:.0f} % | {act} / {tot} | {fps:.0f} | {fps / g:.2f}x |", flush=True)


if __name__ == "__main__":
    main()
