"""oracle/make_golden_gpu.py -- TEST INFRASTRUCTURE.  Run ON THE B200 BOX:

    gpurun -- 'python oracle/make_golden_gpu.py'        (writes gpurun_out/golden/*)

Runs the UNMODIFIED reference on the GPU: its CUDA C kernel strings are compiled by NVRTC for
sm_100a and launched through a minimal ``cupy`` shim (oracle/ref_env.py), everything else is the
reference's own Python on torch/cuDNN.  Needs baseline/_ref (staged by __graft_entry__.build()).
The files it writes are copied to tests/golden/ and committed:

  ref_kernels_<case>.npz       inputs + outputs of split / combine / transfer / repad kernels
  swiftnet_gpu_fp16_clip.pt    SwiftNet-RN18 + BlockCopyModel, fp16, 512x1024, grid 4x8, 6 frames
  swiftnet_gpu_fp16_full.npz   the BENCHMARKED configuration: 1024x2048, grid 8x16, 30 frames, frame 0 all blocks
                               then 40 of 128 (the masks bench.py uses): argmax of every frame + strided logits
                               + strided frame_state                                              ("full")
  det_ig_kat.npz               InformationGainObjectDetection on seeded box lists                          ("det")
  swiftnet_gpu_fp16_smoke.pt   256x512, 64-px blocks, 3 frames: what __graft_entry__.smoke() compares with ("full")
  reference_timing.json        fps of the reference BlockCopy path on this B200
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_env  # noqa: E402

ref = ref_env.import_reference("gpu")
sys.path.append(os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))  # consumers/ only
from consumers.clips import PolicyFixedFraction, PolicyReplay, deterministic_init_, synthetic_clip  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)
dev = "cuda"


def settings(**kw):
    s = dict(block_policy="all", block_num_classes=19, block_optim_lr=1e-4, block_optim_wd=1e-3,
             block_optim_momentum=0, block_target=0.5, block_complexity_weight=5, block_size=128,
             block_train_interval=4, block_cost_momentum=0.9, block_policy_verbose=False)
    s.update(kw)
    return s


def kernel_goldens():
    from blockcopy.core.tensorwrapper import get_grid_mappings
    from blockcopy.utils.block_funcs import CombineFunction, SplitFunction, TransferFunction
    from blockcopy.utils.blockpad import pad as ref_pad

    cases = [  # name, N, C, GH, GW, BS, pad, dtype, frac
        ("a", 1, 4, 3, 4, 8, 1, torch.float16, 0.4),
        ("b", 2, 6, 2, 3, 4, 1, torch.float16, 0.5),
        ("c", 1, 3, 2, 2, 16, 3, torch.float16, 0.5),
        ("d", 1, 8, 3, 3, 4, 2, torch.float32, 0.6),
        ("e", 2, 5, 2, 4, 2, 1, torch.float32, 0.3),
        ("f", 1, 16, 4, 4, 32, 1, torch.float16, 0.3),
    ]
    for name, N, C, GH, GW, BS, pad, dt, frac in cases:
        g = torch.Generator().manual_seed(hash(name) % 1000)
        H, W = GH * BS, GW * BS
        image = torch.randn(N, C, H, W, generator=g).to(dt)
        grid = torch.rand(N, 1, GH, GW, generator=g) < frac
        prev_grid = torch.rand(N, 1, GH, GW, generator=g) < 0.5
        G = grid.numel()

        def maps(gr):
            ne = int(gr.sum())
            gi, me = get_grid_mappings(ne, gr, ~gr, G, G - ne)
            return gi.int(), me.int()

        gi, me = maps(grid)
        pgi, pme = maps(prev_grid)
        ti = pgi[~grid].int()
        E, Ep = me.numel(), pme.numel()
        split = SplitFunction.apply(torch.empty(E, C, BS, BS, dtype=dt, device=dev), image.to(dev), me.to(dev), gi.to(dev))
        tiles_in = torch.randn(E, C, BS, BS, generator=g).to(dt)
        combine_base = torch.randn(N, C, H, W, generator=g).to(dt)
        combine = CombineFunction.apply(tiles_in.to(dev), combine_base.to(dev).clone(), gi.to(dev), me.to(dev))
        prev_exec = torch.randn(Ep, C, BS, BS, generator=g).to(dt)
        prev_transfer = torch.randn(G - Ep, C, BS, BS, generator=g).to(dt)
        transfer_base = torch.randn(G - E, C, BS, BS, generator=g).to(dt)
        transfer = TransferFunction.apply(transfer_base.to(dev).clone(), prev_exec.to(dev), prev_transfer.to(dev),
                                          pgi.to(dev), ti.to(dev), pad)
        repad = ref_pad(tiles_in.to(dev), transfer, gi.to(dev), me.to(dev), pad)
        torch.cuda.synchronize()
        np.savez_compressed(
            os.path.join(OUT, f"ref_kernels_{name}.npz"), dtype=str(dt).split(".")[-1], BS=BS, pad=pad,
            image=image.numpy(), grid=grid.numpy(), grid_idx=gi.numpy(), mapping_exec=me.numpy(),
            split=split.cpu().numpy(), tiles_in=tiles_in.numpy(), combine_base=combine_base.numpy(),
            combine=combine.cpu().numpy(), prev_exec=prev_exec.numpy(), prev_transfer=prev_transfer.numpy(),
            transfer_idx=ti.numpy(), transfer_base=transfer_base.numpy(), transfer=transfer.cpu().numpy(),
            repad=repad.cpu().numpy())
        print("kernel golden", name, "E", E, "T", G - E)


def build_reference_swiftnet(init_seed, gain, BS, policy="all"):
    from lib.models.swiftnet.backbones.resnet import resnet18
    from lib.models.swiftnet.swiftnet import SwiftNet
    from lib.utils import bn_fusion

    net = SwiftNet(resnet18(pretrained=False), num_classes=19, num_features=128, use_spp=True).eval()
    deterministic_init_(net, seed=init_seed, gain=gain)
    model = ref.BlockCopyModel(net, settings(block_size=BS, block_policy=policy)).eval().to(dev)
    model = bn_fusion.fuse_bn_recursively(model)
    model = model.half()
    if model.policy.net is not None:
        model.policy.net = model.policy.net.float()
    return model


def swiftnet_fp16_clip():
    H, W, BS, T, gain = 512, 1024, 128, 6, 0.8
    torch.manual_seed(0)
    random.seed(0)
    model = build_reference_swiftnet(0, gain, BS)
    g = torch.Generator().manual_seed(1)
    grids = [torch.ones(1, 1, H // BS, W // BS, dtype=torch.bool)]
    for frac in (0.3, 0.5, 0.0, 0.25, 1.0):
        cells = grids[0].numel()
        m = torch.zeros(cells, dtype=torch.bool)
        m[torch.randperm(cells, generator=g)[: round(frac * cells)]] = True
        grids.append(m.view_as(grids[0]))
    model.policy = PolicyReplay(BS, grids)
    clip = synthetic_clip(T, H, W, seed=3, dtype=torch.float16, device=dev)
    outs, states = [], []
    with torch.no_grad():
        model.reset_temporal()
        for t in range(T):
            out = model(clip[t])
            assert torch.isfinite(out).all(), "fp16 overflow in the fixture: lower the init gain"
            outs.append(out.clone())
            states.append(model.policy_meta["frame_state"].clone())
    fix = dict(H=H, W=W, BS=BS, T=T, clip_seed=3, init_seed=0, init_gain=gain,
               grids=torch.stack(grids).to(torch.uint8),
               argmax=torch.stack([o.argmax(1).to(torch.uint8).cpu() for o in outs]),
               logits_strided=torch.stack([o[:, :, ::4, ::4].clone().cpu() for o in outs]),
               frame_state_strided=torch.stack([s[:, :, ::8, ::8].clone().cpu() for s in states]),
               logits_abs_mean=[float(o.float().abs().mean()) for o in outs],
               logits_abs_max=[float(o.float().abs().max()) for o in outs])
    torch.save(fix, os.path.join(OUT, "swiftnet_gpu_fp16_clip.pt"))
    print("swiftnet fp16 clip: |logits| mean", [round(v, 3) for v in fix["logits_abs_mean"]], "max",
          [round(v, 2) for v in fix["logits_abs_max"]])


def _run_fixed_fraction_clip(H, W, BS, T, gain, quantize, clip_seed, mask_seed=0):
    torch.manual_seed(0)
    random.seed(0)
    model = build_reference_swiftnet(0, gain, BS)
    model.policy = PolicyFixedFraction(BS, fraction=0.3, quantize=quantize, seed=mask_seed)
    clip = synthetic_clip(T, H, W, seed=clip_seed, dtype=torch.float16, device=dev)
    outs, states, grids = [], [], []
    with torch.no_grad():
        model.reset_temporal()
        for t in range(T):
            out = model(clip[t])
            assert torch.isfinite(out).all(), "fp16 overflow in the fixture: lower the init gain"
            outs.append(out.clone().cpu())
            states.append(model.policy_meta["frame_state"].clone().cpu())
            grids.append(model.policy_meta["grid"].clone().cpu())
    return outs, states, grids


def swiftnet_fp16_full():
    """The configuration bench.py times (BASELINE configs[2]): 1024x2048, BS 128, 30 frames, PolicyFixedFraction
    seed 0 (frame 0 all blocks, then 40 of 128).  Frame t stores the logits / frame_state on a stride-8 / stride-16
    lattice whose offset moves with t, so that over the clip every residue is looked at."""
    H, W, BS, T, gain = 1024, 2048, 128, 30, 0.8
    outs, states, grids = _run_fixed_fraction_clip(H, W, BS, T, gain, quantize=8, clip_seed=0)
    ls = np.stack([o[0, :, (t % 8)::8, ((3 * t) % 8)::8].numpy() for t, o in enumerate(outs)])
    fs = np.stack([s[0, :, (t % 16)::16, ((5 * t) % 16)::16].numpy() for t, s in enumerate(states)])
    np.savez_compressed(
        os.path.join(OUT, "swiftnet_gpu_fp16_full.npz"), H=H, W=W, BS=BS, T=T, clip_seed=0, init_seed=0, init_gain=gain,
        mask_seed=0, grids=np.stack([g.numpy() for g in grids]).astype(np.uint8),
        argmax=np.stack([o.argmax(1)[0].to(torch.uint8).numpy() for o in outs]), logits_strided=ls, frame_state_strided=fs,
        logits_abs_max=np.array([float(o.float().abs().max()) for o in outs]),
        logits_abs_mean=np.array([float(o.float().abs().mean()) for o in outs]))
    print("swiftnet fp16 full clip: exec", [int(g.sum()) for g in grids][:4], "... |logits| max",
          round(max(float(o.float().abs().max()) for o in outs), 2))
    # the small clip smoke() runs
    H, W, BS, T = 256, 512, 64, 3
    outs, states, grids = _run_fixed_fraction_clip(H, W, BS, T, gain, quantize=2, clip_seed=0)
    torch.save(dict(H=H, W=W, BS=BS, T=T, clip_seed=0, init_seed=0, init_gain=gain, mask_seed=0, quantize=2,
                    grids=torch.stack(grids).to(torch.uint8), logits=torch.stack(outs),
                    argmax=torch.stack([o.argmax(1).to(torch.uint8) for o in outs]),
                    frame_state_last=states[-1]), os.path.join(OUT, "swiftnet_gpu_fp16_smoke.pt"))
    print("swiftnet fp16 smoke clip: exec", [int(g.sum()) for g in grids])


def detection_information_gain():
    """InformationGainObjectDetection of the UNMODIFIED reference (policy/information_gain.py:43-108; its IoU-gain
    mask is hard-wired to device 'cuda', so this fixture needs the GPU box) on seeded random box lists shaped like
    mmdet's bbox_results (one class, batch 1) -> det_ig_kat.npz."""
    from blockcopy.policy.information_gain import InformationGainObjectDetection

    rng = np.random.RandomState(7)
    H, W = 256, 512

    def boxes(n, jitter_of=None):
        if jitter_of is not None and len(jitter_of):
            b = jitter_of[:n].copy()
            b[:, :4] += rng.uniform(-6, 6, size=(len(b), 4)).astype(np.float32)
        else:
            b = np.zeros((0, 5), np.float32)
        extra = n - len(b)
        if extra > 0:
            x1 = rng.uniform(0, W - 40, extra); y1 = rng.uniform(0, H - 40, extra)
            w = rng.uniform(8, 120, extra); h = rng.uniform(8, 90, extra)
            e = np.stack([x1, y1, x1 + w, y1 + h, rng.uniform(0.05, 1.0, extra)], 1).astype(np.float32)
            b = np.concatenate([b, e], 0)
        b[:, 4] = rng.uniform(0.05, 1.0, len(b)).astype(np.float32)
        return b.astype(np.float32)

    ig = InformationGainObjectDetection(num_classes=1)
    frames = [boxes(12)]
    frames.append(boxes(14, frames[0]))
    frames.append(boxes(0))                      # nothing detected
    frames.append(boxes(9))
    f4 = boxes(11, frames[3])
    f4[0, :4] = (W - 30.5, H - 20.5, W + 25.0, H + 9.0)   # sticks out of the frame: slices clip
    f4[1, :4] = (-9.0, 40.0, 31.0, 90.0)                  # negative start: Python slicing counts from the end
    frames.append(f4)
    out = dict(H=H, W=W, n_frames=len(frames))
    inputs = torch.zeros(1, 3, H, W, device=dev)
    for t, f in enumerate(frames):
        out[f"boxes_{t}"] = f
        meta = dict(inputs=inputs, outputs=[[f]], outputs_prev=[[frames[t - 1]]] if t else None)
        out[f"repr_{t}"] = ig.get_output_repr(meta).cpu().numpy()
        if t:
            out[f"gain_{t}"] = ig(meta).cpu().numpy()
    np.savez_compressed(os.path.join(OUT, "det_ig_kat.npz"), **out)
    print("detection information gain fixture:", len(frames), "frames")


def time_reference(H=1024, W=2048, clips=3, T=30, fraction=0.3):
    """fps of the reference BlockCopy path (baseline 2): SwiftNet-RN18 fp16,
    random init, seeded ~30 % masks (frame 0 all blocks), cudnn.benchmark on, timings level 0."""
    torch.backends.cudnn.benchmark = True
    model = build_reference_swiftnet(0, 0.8, 128)
    model.policy = PolicyFixedFraction(128, fraction=fraction, quantize=8, seed=0)
    clip = synthetic_clip(T, H, W, seed=0, dtype=torch.float16, device=dev)

    def run_clip():
        model.reset_temporal()
        with torch.no_grad():
            for f in clip:
                out = model(f)
        return out

    run_clip()  # warm-up: NVRTC compiles, cudnn.benchmark
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(clips):
        run_clip()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # dense cuDNN fp16 forward of the same model for comparison (baseline 3)
    dense = model.base_model
    with torch.no_grad():
        for _ in range(3):
            dense(clip[0])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(20):
            dense(clip[0])
        torch.cuda.synchronize()
        dense_dt = (time.perf_counter() - t1) / 20
    res = dict(reference_blockcopy_fps=clips * T / dt, ms_per_frame=1000 * dt / (clips * T), frames=clips * T,
               H=H, W=W, fraction=fraction, num_exec=model.policy.num_exec_for(128),
               dense_cudnn_fp16_fps=1 / dense_dt, gpu=torch.cuda.get_device_name(0),
               note="reference package unmodified; cupy -> NVRTC shim; python (not -O); timings level 0")
    with open(os.path.join(OUT, "reference_timing.json"), "w") as f:
        json.dump(res, f, indent=1)
    print("reference timing:", json.dumps(res))


def time_reference_sweep(out_path=None):
    """BASELINE config 5: the reference BlockCopy path (same shim, same settings as time_reference) over the
    active-fraction sweep at 1024x2048 and 2048x4096; one warm-up clip + five timed 30-frame clips per point
    (median clip reported: the path is host-bound and noisy)."""
    torch.backends.cudnn.benchmark = True
    rows = []
    for H, W in ((1024, 2048), (2048, 4096)):
        clip = synthetic_clip(30, H, W, seed=0, dtype=torch.float16, device=dev)
        for fraction in (0.05, 0.10, 0.20, 0.30, 0.50, 0.75, 1.00):
            model = build_reference_swiftnet(0, 0.8, 128)
            model.policy = PolicyFixedFraction(128, fraction=fraction, quantize=8, seed=0)

            def run_clip():
                model.reset_temporal()
                with torch.no_grad():
                    for f in clip:
                        model(f)

            run_clip()  # warm-up: NVRTC compiles, cudnn.benchmark picks algorithms
            torch.cuda.synchronize()
            per_clip = []
            for _ in range(5):
                t0 = time.perf_counter()
                run_clip()
                torch.cuda.synchronize()
                per_clip.append(30 / (time.perf_counter() - t0))
            per_clip.sort()
            G = (H // 128) * (W // 128)
            rows.append(dict(H=H, W=W, fraction=fraction, num_exec=model.policy.num_exec_for(G), total=G,
                             reference_blockcopy_fps=per_clip[2], best_clip_fps=per_clip[-1], worst_clip_fps=per_clip[0]))
            print("reference sweep:", json.dumps(rows[-1]), flush=True)
            del model
            torch.cuda.empty_cache()
    out_path = out_path or os.path.join(OUT, "reference_sweep.json")
    with open(out_path, "w") as f:
        json.dump(dict(gpu=torch.cuda.get_device_name(0), rows=rows,
                       note="reference package unmodified; cupy -> NVRTC shim; timings level 0"), f, indent=1)


if __name__ == "__main__":
    what = sys.argv[1:] or ["kernels", "clip", "time"]
    if "sweep" in what:
        time_reference_sweep(os.environ.get("BC_SWEEP_OUT"))
    if "kernels" in what:
        kernel_goldens()
    if "clip" in what:
        swiftnet_fp16_clip()
    if "full" in what:
        swiftnet_fp16_full()
    if "det" in what:
        detection_information_gain()
    if "time" in what:
        time_reference()
