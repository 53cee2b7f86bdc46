"""baseline/time_reference_gpu.py -- REFERENCE ARM on the GPU (baseline 2).

    python -O baseline/time_reference_gpu.py [--policy fixed|rl_semseg] [--clips 5] [--frames 30]

Times the UNMODIFIED reference `blockcopy` package (staged copy in baseline/_ref, or /root/reference) driving the
reference's own SwiftNet-RN18 on one GPU: its CUDA C kernel strings are compiled by NVRTC for the device's
architecture and launched through the cupy shim of baseline/ref_env.py; everything else is the reference's
Python on torch / cuDNN.  Same workload as bench.py's own arm: random-init SwiftNet-RN18 (fp16, BN fused by the
reference's bn_fusion), synthetic 30-frame 1024x2048 clip, frame 0 all blocks then 40 of 128 seeded masks
(`fixed`), or the reference's own `rl_semseg` policy with target 0.3 and online training every 3rd frame.  Loop
shape follows semantic_segmentation/test_swiftnet.py:147-197: cudnn.benchmark=True during the warm-up clip, False
while timing; timings level 0; wall clock between two device synchronisations.  Prints ONE JSON line.

Runs in its own interpreter (the reference package is also called `blockcopy`); nothing of this repo's product
(package, kernels, .so) is imported -- only the synthetic clip / mask helpers of consumers/clips.py, which subclass
the REFERENCE's Policy here.
"""
import argparse
import json
import os
import statistics
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--policy", default="fixed", choices=["fixed", "rl_semseg"])
    ap.add_argument("--clips", type=int, default=5)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--fraction", type=float, default=0.3)
    ap.add_argument("--level10", action="store_true", help="one extra clip at timings level 10 (per-region breakdown)")
    args = ap.parse_args()

    import contextlib
    import io
    import random

    import torch

    from baseline import ref_env

    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_env.import_reference("gpu")
    sys.path.append(os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))  # consumers/ only
    from consumers.clips import PolicyFixedFraction, deterministic_init_, synthetic_clip

    assert "baseline/_ref" in ref.__file__ or "/root/reference" in ref.__file__, ref.__file__
    dev = "cuda"
    torch.manual_seed(0)
    random.seed(0)
    settings = dict(block_policy="all" if args.policy == "fixed" else "rl_semseg", block_num_classes=19,
                    block_optim_lr=1e-4, block_optim_wd=1e-3, block_optim_momentum=0, block_target=args.fraction,
                    block_complexity_weight=5, block_size=128, block_train_interval=3, block_cost_momentum=0.9,
                    block_policy_verbose=False)
    with contextlib.redirect_stdout(io.StringIO()):
        from lib.models.swiftnet.backbones.resnet import resnet18
        from lib.models.swiftnet.swiftnet import SwiftNet
        from lib.utils import bn_fusion

        net = SwiftNet(resnet18(pretrained=False), num_classes=19, num_features=128, use_spp=True).eval()
        deterministic_init_(net, seed=0, gain=0.8)
        model = ref.BlockCopyModel(net, settings).eval().to(dev)
        model = bn_fusion.fuse_bn_recursively(model)
    model = model.half()
    if model.policy.net is not None:
        model.policy.net = model.policy.net.float()  # test_swiftnet.py:118-123
    if args.policy == "fixed":
        model.policy = PolicyFixedFraction(128, fraction=args.fraction, quantize=8, seed=0)
    clip = synthetic_clip(args.frames, args.height, args.width, seed=0, dtype=torch.float16, device=dev)

    execd = []

    def run_clip():
        model.reset_temporal()
        with torch.no_grad():
            for f in clip:
                model(f)
                execd.append(int(model.policy_meta["num_exec"]))

    torch.backends.cudnn.benchmark = True
    with contextlib.redirect_stdout(io.StringIO()):
        run_clip()  # warm-up: NVRTC compiles, cudnn.benchmark picks algorithms
        if args.policy == "rl_semseg":
            run_clip()
    torch.cuda.synchronize()
    torch.backends.cudnn.benchmark = False
    del execd[:]
    per_clip = []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(args.clips):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run_clip()
            torch.cuda.synchronize()
            per_clip.append(args.frames / (time.perf_counter() - t0))
    res = dict(impl="reference_gpu_path", policy=args.policy, value=statistics.median(per_clip), unit="frames/s",
               per_clip=per_clip, clips=args.clips, frames_per_clip=args.frames, height=args.height, width=args.width,
               mean_exec_blocks=sum(execd) / max(1, len(execd)), total_blocks=(args.height // 128) * (args.width // 128),
               optimized_python=not __debug__, gpu=torch.cuda.get_device_name(0),
               note="reference blockcopy package + reference SwiftNet, unmodified; cupy -> NVRTC shim; timings level 0; "
                    "cudnn.benchmark on in warm-up, off while timing; median clip")
    if args.level10:
        from blockcopy.utils.profiler import timings
        timings.set_level(10)
        timings.reset()
        timings.add_cnt(args.frames)
        with contextlib.redirect_stdout(io.StringIO()):
            run_clip()
        res["timings_level10"] = repr(timings)
        timings.set_level(0)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
