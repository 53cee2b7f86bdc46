#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: frames/s of SwiftNet-RN18 + BlockCopy at 1024x2048 with
~30 % active blocks on B200(s), plus the HBM roofline of the dominant block kernel and the
reference's CPU path timed beside it.  Prints ONE JSON line (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo (CUDA, C ABI)
    python bench.py --impl reference ...                         # reference arm: dense SwiftNet, CPU torch
    python bench.py --microbench                                 # block kernels only (ncu target)

A "step" is one frame of every stream the rank owns, through blockcopy.BlockCopyModel (policy ->
gather -> block-sparse SwiftNet -> combine).  Streams are independent; ranks never communicate
on the data path (weak scaling: --streams-per-gpu is fixed).  Frames cycle through a seeded
30-frame synthetic clip; frame 0 of each clip executes every block (API contract), the others
exactly 40 of 128 blocks (30 % rounded up to the reference's multiple of 8), masks seeded and
generated on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200")
sys.path.insert(0, PKG)

import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=90)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--fraction", type=float, default=0.3)
    ap.add_argument("--clip-length", type=int, default=30)
    ap.add_argument("--streams-per-gpu", type=int, default=1)
    ap.add_argument("--batch", type=int, default=1, help="streams batched along N in one model call "
                    "(the reference's speed configs use --batch-size 2)")
    ap.add_argument("--policy", default="fixed", choices=["fixed", "rl_semseg"])
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA graphs")
    ap.add_argument("--microbench", action="store_true", help="only the block-kernel microbenchmarks")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--concurrent-streams", action="store_true",
                    help="with --streams-per-gpu > 1: one CUDA stream per video stream (frames overlap on the GPU)")
    ap.add_argument("--shared-policy", action="store_true",
                    help="rl_semseg only: one policy for all ranks (all-reduce of the policy gradients over NCCL)")
    ap.add_argument("--skip-batched", action="store_true", help="skip the extra batch-8 throughput measurement")
    ap.add_argument("--repeats", type=int, default=20, help="timed windows of --steps steps each; the median is reported")
    ap.add_argument("--skip-config4", action="store_true", help="skip the 64-stream (config 4) sub-record")
    ap.add_argument("--skip-rl", action="store_true", help="skip the rl_semseg sub-record")
    ap.add_argument("--skip-microbench", action="store_true")
    ap.add_argument("--skip-reference-gpu", action="store_true", help="do not time the reference's BlockCopy path on the GPU")
    return ap.parse_args()


# =============================================================================================== helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return world, rank, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()


# =============================================================================================== workload
_ORIG_AFFINITY = {}


def pin_to_local_numa(local_rank: int):
    """Restrict this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is
    allocated (first touch => the staging buffers live in node-local memory; round 1's 8-GPU end-to-end numbers
    were limited by every rank's pinned pool sitting on one node).  Returns a dict for the JSON line."""
    info = {"numa_node": None, "cpus": None}
    bus = None
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:
        bus = None
    if bus is None:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bus = out.splitlines()[0].strip() if out else None
        except Exception:
            bus = None
    try:
        dom, rest = bus.split(":", 1)
        dev = f"{int(dom, 16):04x}:{rest.lower()}"
        base = f"/sys/bus/pci/devices/{dev}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = ids & set(os.sched_getaffinity(0))
        info = {"numa_node": node, "cpus": len(allowed), "pci": dev}
        if node >= 0 and allowed:
            _ORIG_AFFINITY.setdefault("cpus", set(os.sched_getaffinity(0)))
            os.sched_setaffinity(0, allowed)
            info["pinned"] = True
    except Exception:
        pass
    return info


def _reference_swiftnet():
    """The reference's OWN SwiftNet files (staged unmodified under baseline/_ref) when present, else None."""
    ref_ss = os.path.join(ROOT, "baseline", "_ref", "semantic_segmentation")
    if not os.path.isdir(os.path.join(ref_ss, "lib", "models", "swiftnet")) or os.environ.get("BLOCKCOPY_BENCH_OWN_MODEL") == "1":
        return None
    import contextlib
    import io

    import blockcopy  # noqa: F401  (this repo's package: the reference model files import its decorator / timings)

    if ref_ss not in sys.path:
        sys.path.insert(0, ref_ss)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from lib.models.swiftnet.backbones.resnet import resnet18
            from lib.models.swiftnet.swiftnet import SwiftNet
            from lib.utils import bn_fusion

            net = SwiftNet(resnet18(pretrained=False), num_classes=19, num_features=128, use_spp=True).eval()
        return net, bn_fusion
    except Exception as e:  # pragma: no cover
        sys.stderr.write(f"bench: reference SwiftNet files not usable ({e}); using consumers/swiftnet_rn18.py\n")
        return None


MODEL_CODE = {"value": None}


def build_model(args, device, policy=None, native_training=True):
    """SwiftNet-RN18 behind this package's BlockCopyModel, set up like the reference driver does
    (test_swiftnet.py:107-123): wrap, BN fusion, .half(), policy net in fp32."""
    import contextlib
    import io

    import blockcopy
    from blockcopy.core.argparser import default_settings
    from consumers.clips import PolicyFixedFraction, deterministic_init_
    from consumers.swiftnet_rn18 import SwiftNetRN18, fuse_conv_bn_

    policy = policy or args.policy
    settings = default_settings(block_policy="rl_semseg" if policy == "rl_semseg" else "all",
                                block_target=args.fraction, block_train_interval=3)
    if not args.no_graphs:
        settings["block_cuda_graphs"] = True
    if getattr(args, "shared_policy", False):
        settings["block_policy_shared"] = True
    if os.environ.get("BLOCKCOPY_POLICY_FUSED", "1") == "0":  # A/B: policy trunk always through torch/cuDNN
        settings["block_policy_fused"] = False
    # training frames of the policy net: this repo's kernels (policy/fused_train.py, the default) or torch autograd over
    # cuDNN graph replays (A/B: native_training=False / BLOCKCOPY_POLICY_FUSED_TRAINING=0)
    settings["block_policy_fused_training"] = bool(native_training) and os.environ.get("BLOCKCOPY_POLICY_FUSED_TRAINING", "1") == "1"
    # the reference driver's order (test_swiftnet.py:104-123): base model in eval mode, THEN wrapped (the policy net
    # stays in train mode, so the BN fusion below leaves its BatchNorms alone), fused, .half(), policy net back to fp32
    ref = _reference_swiftnet()
    if ref is not None:
        net, bn_fusion = ref
        deterministic_init_(net, seed=0, gain=0.8)
        model = blockcopy.BlockCopyModel(net.eval(), settings).to(device)
        with contextlib.redirect_stdout(io.StringIO()):
            model = bn_fusion.fuse_bn_recursively(model)
        MODEL_CODE["value"] = "reference's own lib/models/swiftnet/*.py + lib/utils/bn_fusion.py (baseline/_ref, unmodified)"
    else:
        net = deterministic_init_(SwiftNetRN18().eval(), seed=0, gain=0.8)
        model = blockcopy.BlockCopyModel(net, settings).to(device)
        fuse_conv_bn_(model)
        MODEL_CODE["value"] = "consumers/swiftnet_rn18.py (architecture- and state_dict-identical; baseline/_ref not staged)"
    model = model.eval()
    model = model.half()
    if policy == "fixed":
        model.policy = PolicyFixedFraction(128, fraction=args.fraction, quantize=8, seed=0)
    else:
        model.policy.net = model.policy.net.float().train()
    return model


def _advance(models, t, clip_len):
    k = t % clip_len
    if k == 0:
        for s, model in enumerate(models):
            model.reset_temporal()
            if hasattr(model.policy, "reseed"):
                model.policy.reseed(1000 * s + t // clip_len)
    return k


_SIDE_STREAMS = {}


def run_frames(models, clips, start, count, clip_len, concurrent=False):
    """Device-resident inputs: advance every stream by `count` frames from global frame `start`.
    concurrent: every video stream issues on its own CUDA stream (forked from / joined to the current one), so
    frames of different video streams overlap on the GPU."""
    out = None
    main = torch.cuda.current_stream()
    side = None
    if concurrent and len(models) > 1:
        side = [_SIDE_STREAMS.setdefault((main.device, s), torch.cuda.Stream(device=main.device)) for s in range(len(models))]
        for st in side:
            st.wait_stream(main)
    with torch.no_grad():
        for t in range(start, start + count):
            k = _advance(models, t, clip_len)
            for s, model in enumerate(models):
                if side is None:
                    out = model(clips[s][k])
                else:
                    with torch.cuda.stream(side[s]):
                        out = model(clips[s][k])
    if side is not None:
        for st in side:
            main.wait_stream(st)
    return out


class HostPipeline:
    """End-to-end loop with HOST buffers, through the public API: every step uploads that step's frames from pinned
    host memory and reads the step's results back to pinned host memory, all inside the timed region.

    u8=True (the `e2e` number): what a video pipeline moves -- the decoded uint8 frame (H,W,3) goes up,
    bc_frame_from_u8 normalises it on the device, the model runs, bc_upsample_argmax turns the logits into the
    full-resolution uint8 label map, which goes down (test_swiftnet.py:187-197 does upload -> model -> interpolate
    -> max -> .cpu()).  u8=False (`e2e_fp16`): the fp16 network input goes up, the fp16 logits come down.

    All S streams of the rank travel in ONE copy per direction and step (one pinned (S,B,...) buffer per step).
    Three CUDA streams: upload (+ normalise), compute, download (label map + copy), double-buffered, so the copies
    and the two driver-side kernels of frame t+1 / t-1 overlap the model of frame t."""

    def __init__(self, models, host_clips, device, u8=True):
        self.models, self.device, self.u8 = models, device, u8
        # uint8 frames as lazily normalised model inputs (blockcopy.U8Frame: normalisation fused into the first gather,
        # bc_blocks_from_u8) save 25 MB of HBM traffic per frame but put ~1.5 us more on the frame's critical path than
        # normalising whole frames on the upload stream: measured +2 % e2e with 8 frames per step, -1.7 % with one
        # (tools/u8_gather_bench.py) -- so: lazy when a step carries more than one frame
        env = os.environ.get("BC_E2E_LAZY_U8")
        self.lazy_u8 = (env == "1") if env is not None else (len(models) * host_clips[0][0].shape[0] > 1)
        S = len(models)
        L = len(host_clips[0])
        B, _, H, W = host_clips[0][0].shape
        self.S, self.B, self.H, self.W = S, B, H, W
        # copies and the two driver-side kernels run at the LOWEST priority, the model on a high-priority stream: the
        # block scheduler hands SMs to the frame's (latency-bound) kernels first, the I/O kernels fill the gaps
        self.up, self.down = torch.cuda.Stream(device=device, priority=0), torch.cuda.Stream(device=device, priority=0)
        self.compute = torch.cuda.Stream(device=device, priority=-1)
        if u8:
            from consumers.frame_io import CITYSCAPES_MEAN, CITYSCAPES_STD, BlockLabelMap, FrameNormalizer
            from blockcopy.core.frame import U8Frame
            self.norm, self.U8Frame = FrameNormalizer(), U8Frame
            mean = torch.tensor(CITYSCAPES_MEAN).view(1, 3, 1, 1)
            std = torch.tensor(CITYSCAPES_STD).view(1, 3, 1, 1)
            # the synthetic fp16 clips as decoded uint8 frames (S,B,H,W,3) per step, pinned (allocated by THIS rank's
            # threads after pin_to_local_numa: node-local)
            self.host_in = []
            for k in range(L):
                buf = torch.empty((S, B, H, W, 3), dtype=torch.uint8).pin_memory()
                for s in range(S):
                    f = host_clips[s][k].float()
                    buf[s] = ((f * std + mean) * 255).round_().clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
                self.host_in.append(buf)
            self.dev_raw = [torch.empty((S, B, H, W, 3), dtype=torch.uint8, device=device) for _ in range(2)]
            self.dev_in = [torch.empty((S, B, 3, H, W), dtype=torch.float16, device=device) for _ in range(2)]
            # ONE persistent device label buffer for all streams, updated in place, block-sparsely, every frame
            # (consumers/frame_io.BlockLabelMap); downloads are ordered behind the updates on the download stream
            self.dev_res = [torch.empty((S, B, H, W), dtype=torch.uint8, device=device)] * 2
            self.label_maps = [BlockLabelMap(out=self.dev_res[0][s]) for s in range(S)]
            self.host_out = [torch.empty((S, B, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
        else:
            self.host_in = []
            for k in range(L):
                buf = torch.empty((S, B, 3, H, W), dtype=torch.float16).pin_memory()
                for s in range(S):
                    buf[s] = host_clips[s][k]
                self.host_in.append(buf)
            self.dev_in = [torch.empty((S, B, 3, H, W), dtype=torch.float16, device=device) for _ in range(2)]
            self.dev_res = [torch.empty((S, B, 19, H // 4, W // 4), dtype=torch.float16, device=device) for _ in range(2)]
            self.host_out = [torch.empty((S, B, 19, H // 4, W // 4), dtype=torch.float16).pin_memory() for _ in range(2)]
        self.h2d_bytes = self.host_in[0].numel() * self.host_in[0].element_size()
        self.d2h_bytes = self.host_out[0].numel() * self.host_out[0].element_size()
        self.in_ready = [torch.cuda.Event() for _ in range(2)]
        self.in_free = [None, None]     # compute finished reading dev_in[slot]
        self.res_free = [None, None]    # download of dev_res[slot] finished
        self.out_read = [[None, None] for _ in range(S)]  # the model's ping-pong output buffer was read out

    SKIP = set(filter(None, os.environ.get("BC_E2E_SKIP", "").split(",")))  # diagnostics: h2d,d2h,norm,labels

    def _upload(self, t, clip_len):
        slot = t % 2
        with torch.cuda.stream(self.up):
            if self.in_free[slot] is not None:
                self.up.wait_event(self.in_free[slot])
            if self.u8:
                if "h2d" not in self.SKIP:
                    self.dev_raw[slot].copy_(self.host_in[t % clip_len], non_blocking=True)
                if "norm" not in self.SKIP and not self.lazy_u8:
                    self.norm(self.dev_raw[slot].view(self.S * self.B, self.H, self.W, 3),
                              out=self.dev_in[slot].view(self.S * self.B, 3, self.H, self.W))
            elif "h2d" not in self.SKIP:
                self.dev_in[slot].copy_(self.host_in[t % clip_len], non_blocking=True)
            self.in_ready[slot].record(self.up)

    def run(self, start, count, clip_len):
        caller = torch.cuda.current_stream()
        main = self.compute
        main.wait_stream(caller)
        with torch.no_grad(), torch.cuda.stream(main):
            self._upload(start, clip_len)
            for t in range(start, start + count):
                _advance(self.models, t, clip_len)
                slot = t % 2
                if t + 1 < start + count:
                    self._upload(t + 1, clip_len)
                main.wait_event(self.in_ready[slot])
                outs = []
                for s, model in enumerate(self.models):
                    if self.out_read[s][slot] is not None:
                        main.wait_event(self.out_read[s][slot])  # the output buffer this call rewrites was read out
                    if self.u8 and self.lazy_u8:
                        # the uint8 frame itself is the model's input: steady frames normalise only the executed blocks,
                        # inside their first gather (bc_blocks_from_u8); dev_in[slot][s] receives the whole normalised
                        # frame only when somebody asks for it (first frame of a clip)
                        outs.append(model(self.U8Frame(self.dev_raw[slot][s], out=self.dev_in[slot][s])))
                    else:
                        outs.append(model(self.dev_in[slot][s]))
                    if self.u8 and "labels" not in self.SKIP:
                        # block-sparse label update right behind the frame, on the compute stream: ~1/3 of the label
                        # map per frame; run concurrently (download stream) its 1000+ small CTAs took the SMs' register
                        # files away from the next frame's first kernels and cost more than they do in line
                        self.label_maps[s].update(outs[-1], model.policy_meta["grid"])
                done = torch.cuda.Event()
                done.record(main)
                self.in_free[slot] = done
                with torch.cuda.stream(self.down):
                    self.down.wait_event(done)
                    if self.res_free[slot] is not None:
                        self.down.wait_event(self.res_free[slot])
                    for s, out in enumerate(outs):
                        if "labels" in self.SKIP:
                            pass
                        elif self.u8:
                            pass  # done on the compute stream (see above)
                        else:
                            self.dev_res[slot][s].copy_(out, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self.down)
                        self.out_read[s][slot] = ev
                    if "d2h" not in self.SKIP:
                        self.host_out[slot].copy_(self.dev_res[slot], non_blocking=True)
                    self.res_free[slot] = torch.cuda.Event()
                    self.res_free[slot].record(self.down)
        caller.wait_stream(main)
        caller.wait_stream(self.down)
        caller.wait_stream(self.up)


def _summary(ms):
    """min / median / p90 / max of a list of window times (the JSON line stays readable)."""
    v = sorted(ms)
    return {"n": len(v), "min": v[0], "median": statistics.median(v), "p90": v[min(len(v) - 1, int(0.9 * len(v)))], "max": v[-1]}


def timed_windows(fn, steps, repeats, world, device):
    """`repeats` back-to-back windows of exactly `steps` steps each; every window is bracketed by a barrier +
    torch.cuda.synchronize() on both sides and timed with CUDA events on the launching stream.  Returns the list of
    per-window milliseconds, MAX over ranks per window.  fn(start_step, steps)."""
    ms = []
    pos = 0
    for _ in range(repeats):
        torch.cuda.synchronize()
        barrier(world)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(pos, steps)
        b.record()
        torch.cuda.synchronize()
        barrier(world)
        ms.append(a.elapsed_time(b))
        pos += steps
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor(ms, dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.tolist()
    return ms


def _device_clip(L, H, W, seed, batch, device):
    """synthetic_clip's recipe generated on the device (config 4 needs 64 streams x 30 frames: too slow on the CPU)."""
    g = torch.Generator(device=device).manual_seed(seed)
    frame = torch.randn(batch, 3, H, W, generator=g, device=device)
    frames = []
    for _ in range(L):
        frames.append(frame.half())
        noise = torch.randn(frame.shape, generator=g, device=device)
        gate = torch.rand(frame.shape, generator=g, device=device) > 0.7
        frame = frame + 0.3 * noise * gate
    return frames


def bench_config4(args, device, world, rank, total_streams=64, group=8, steps=30):
    """BASELINE configs[3]: 64 independent 1024x2048 streams sharded over the ranks (stream_id % world), the
    64/world streams of a GPU batched along N in wrappers of `group` streams (grid (N,1,GH,GW), mapping_exec spans the
    batch).  Device-resident frames; whole-job frames/s, max over ranks."""
    H, W, L = args.height, args.width, args.clip_length
    per_rank = total_streams // world
    group = min(group, per_rank)
    wrappers = per_rank // group
    models = [build_model(args, device, policy="fixed") for _ in range(wrappers)]
    clips = [_device_clip(L, H, W, seed=10_000 + 100 * rank + w, batch=group, device=device) for w in range(wrappers)]
    run_frames(models, clips, 0, 2 * L + 2, L)
    ms = timed_windows(lambda pos, n: run_frames(models, clips, 2 + pos, n, L), steps, 3, world, device)
    med = statistics.median(ms)
    frames = float(steps * per_rank * world)
    del models, clips
    torch.cuda.empty_cache()
    return {"streams_total": per_rank * world, "streams_per_gpu": per_rank, "batched_along_N": group,
            "wrappers_per_gpu": wrappers, "value": frames / (med * 1e-3), "unit": "frames/s", "steps": steps,
            "ms_per_step": med / steps, "windows_ms": _summary(ms), "data": "synthetic, device-resident"}


def bench_rl(args, device, steps=90, native_training=True):
    """The mode the reference ships (`--block-policy rl_semseg`): policy net fp32, Bernoulli sampling, online
    REINFORCE step every 3rd frame (block_train_interval 3), target 0.3.  Single stream, device-resident frames."""
    from consumers.clips import synthetic_clip

    H, W, L = args.height, args.width, args.clip_length
    import random

    random.seed(0)
    torch.manual_seed(0)
    m = build_model(args, device, policy="rl_semseg", native_training=native_training)
    clip = [f.to(device) for f in synthetic_clip(L, H, W, seed=3, dtype=torch.float16)]
    run_frames([m], [clip], 0, 3 * L, L)
    execd = []

    def fn(pos, n):
        with torch.no_grad():
            for t in range(pos, pos + n):
                k = _advance([m], t, L)
                m(clip[k])
                execd.append(m.policy_meta["num_exec"])

    ms = timed_windows(fn, steps, 5, 1, device)
    med = statistics.median(ms)
    res = {"policy": "rl_semseg", "value": steps / (med * 1e-3), "unit": "frames/s", "steps": steps, "windows_ms": _summary(ms),
           "mean_exec_blocks": sum(execd) / len(execd), "total_blocks": (H // 128) * (W // 128),
           "train_interval": 3, "policy_training_frames": "native kernels (policy/fused_train.py)" if native_training
           else "torch autograd over cuDNN graph replays",
           "note": "random-init policy net trained online; masks are sampled (not replayed), so "
                                        "the executed fraction follows the policy, see mean_exec_blocks; the fused "
                                        "policy trunk runs fp16 operands (reference: fp32), masks are statistically, "
                                        "not bitwise, equivalent"}
    del m, clip
    torch.cuda.empty_cache()
    return res


def reference_gpu_live(policy="fixed", clips=5, timeout=600):
    """BASELINE.md section 3, baseline 2, measured IN THIS RUN: the unmodified reference package + reference
    SwiftNet on this GPU through the NVRTC cupy shim, `python -O`, in its own interpreter
    (baseline/time_reference_gpu.py)."""
    script = os.path.join(ROOT, "baseline", "time_reference_gpu.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "blockcopy")) and not os.path.isdir("/root/reference/blockcopy"):
        return {"unavailable": "reference not staged (baseline/_ref)"}
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        res = subprocess.run([sys.executable, "-O", script, "--policy", policy, "--clips", str(clips)], capture_output=True,
                             text=True, timeout=timeout, env=env, cwd=ROOT)
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not line:
            return {"unavailable": f"rc {res.returncode}: {res.stderr.strip()[-300:]}"}
        out = json.loads(line[-1])
        out["measured_in_this_run"] = True
        return out
    except Exception as e:  # pragma: no cover
        return {"unavailable": repr(e)}


def bench_ours(args):
    from blockcopy import _C
    from consumers.clips import synthetic_clip

    world, rank, local = dist_setup(args)
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = pin_to_local_numa(local)
    torch.backends.cudnn.benchmark = True
    peaks = load_peaks()
    H, W, L, S, B = args.height, args.width, args.clip_length, args.streams_per_gpu, args.batch
    K, R = args.steps, args.repeats

    models = [build_model(args, device) for _ in range(S)]
    host_clips = [synthetic_clip(L, H, W, seed=100 * rank + s, batch=B, dtype=torch.float16) for s in range(S)]
    clips = [[f.to(device) for f in hc] for hc in host_clips]
    G = (H // 128) * (W // 128)
    num_exec = models[0].policy.num_exec_for(G) if hasattr(models[0].policy, "num_exec_for") else None

    # ---- setup (not steps): two full clips so that every CUDA graph (one per block count) is captured,
    #      cuDNN has picked its algorithms and all planes exist; then the W warm-up steps
    conc = bool(args.concurrent_streams)
    run_frames(models, clips, 0, 2 * L, L, conc)
    # ---- device-resident throughput ("value"): R windows of K steps, median window ------------------------------
    run_frames(models, clips, 0, args.warmup, L, conc)
    torch.cuda.synchronize()
    # enough windows for >= ~2 s of timed work (a 20-step window is only ~7 ms; the clock sampler needs samples)
    t_est = time.perf_counter()
    run_frames(models, clips, args.warmup, K, L, conc)
    torch.cuda.synchronize()
    t_est = time.perf_counter() - t_est
    R = int(min(400, max(R, 2.0 / max(t_est, 1e-4))))
    if world > 1:
        import torch.distributed as dist

        r_t = torch.tensor([R], device=device)
        dist.all_reduce(r_t, op=dist.ReduceOp.MAX)
        R = int(r_t.item())
    n0 = _C.launch_count()
    with ClockSampler(local) as clk:
        win = timed_windows(lambda pos, n: run_frames(models, clips, args.warmup + pos, n, L, conc), K, R, world, device)
    launches = (_C.launch_count() - n0) // R
    med = statistics.median(win)
    frames_step = float(S * B * world)
    value = K * frames_step / (med * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e") ---------------------------------
    e2e, e2e_fp16 = None, None
    if not args.skip_e2e:
        Re = max(5, R // 2)
        pipe = HostPipeline(models, host_clips, device, u8=True)
        pipe.run(0, min(args.warmup, L), L)
        w8 = timed_windows(lambda pos, n: pipe.run(args.warmup + pos, n, L), K, Re, world, device)
        e2e = {"value": K * frames_step / (statistics.median(w8) * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes, "windows_ms": _summary(w8),
               "note": "pinned uint8 (H,W,3) frames of all the rank's streams -> ONE H2D copy -> " + (
                   "model(U8Frame): normalisation inside the first gather of the executed blocks (bc_blocks_from_u8; whole "
                   "frames only on a clip's first frame)" if pipe.lazy_u8 else "bc_frame_from_u8 on the upload stream -> model()") +
                       " -> block-sparse bc_upsample_argmax_blocks -> ONE D2H copy of the full-resolution uint8 label maps; upload, "
                       "compute and download on three streams, double-buffered; pinned buffers NUMA-local",
               "numa": numa}
        del pipe
        pipe = HostPipeline(models, host_clips, device, u8=False)
        pipe.run(0, min(args.warmup, L), L)
        w16 = timed_windows(lambda pos, n: pipe.run(args.warmup + pos, n, L), K, 3, world, device)
        e2e_fp16 = {"value": K * frames_step / (statistics.median(w16) * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "note": "the interface the reference's driver uses on the way up: fp16 network input up, fp16 logits down"}
        del pipe
    del clips, host_clips
    models.clear()
    torch.cuda.empty_cache()

    # ---- secondary records ---------------------------------------------------------------------------------------
    batched = None
    if not args.skip_batched and B == 1 and S == 1 and world == 1:
        B8 = 8
        m8 = build_model(args, device)
        clip8 = _device_clip(L, H, W, seed=7, batch=B8, device=device)
        run_frames([m8], [clip8], 0, 2 * L + 2, L)
        wb = timed_windows(lambda pos, n: run_frames([m8], [clip8], 2 + pos, n, L), L, 3, 1, device)
        batched = {"batch": B8, "value": B8 * L / (statistics.median(wb) * 1e-3), "unit": "frames/s",
                   "ms_per_step": statistics.median(wb) / L, "steps": L}
        del m8, clip8
        torch.cuda.empty_cache()
    config4 = None
    if not args.skip_config4 and B == 1 and S == 1 and 64 % world == 0:
        config4 = bench_config4(args, device, world, rank)
    rl = None
    if not args.skip_rl and args.policy == "fixed" and world == 1:
        rl = bench_rl(args, device)
        alt = bench_rl(args, device, native_training=False)
        rl["torch_autograd_training"] = {k: alt[k] for k in ("value", "unit", "mean_exec_blocks", "policy_training_frames")}

    # ---- kernels: roofline of the dominant block kernel + the others ----------------------------------
    kern = microbench(device, peaks) if rank == 0 and not args.skip_microbench else None
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cpu = cpu_dense_baseline(H, W, frames=5)
    ref_gpu, ref_gpu_rl = None, None
    if rank == 0 and world == 1 and not args.skip_reference_gpu:
        torch.cuda.empty_cache()
        ref_gpu = reference_gpu_live("fixed")
        if "value" in ref_gpu:
            ref_gpu["ratio_ours_over_reference"] = value / ref_gpu["value"]
            if ref_gpu.get("per_clip"):  # the reference's clips scatter on some boxes (host-bound): also against its best clip
                ref_gpu["ratio_ours_over_reference_best_clip"] = value / max(ref_gpu["per_clip"])
        if rl is not None:
            ref_gpu_rl = reference_gpu_live("rl_semseg", clips=3)
            if "value" in ref_gpu_rl:
                ref_gpu_rl["ratio_ours_over_reference"] = rl["value"] / ref_gpu_rl["value"]
                ref_gpu_rl["caveat"] = "each side's own random-init policy decides how many blocks run (mean_exec_blocks)"

    if rank == 0:
        line = {
            "metric": "frames/s @1024x2048, 30% active blocks (SwiftNet-RN18 + BlockCopy)", "value": value,
            "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": med / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "timing": {"windows": R, "window_ms": _summary(win), "statistic": "median window of `steps` steps, each window bracketed "
                       "by barrier + synchronize, CUDA events, max over ranks per window"},
            "config": {"workload": "configs[2]: SwiftNet-RN18 + BlockCopy, synthetic 1024x2048 30-frame clips, "
                                   "random-init weights, seeded masks",
                       "model_code": MODEL_CODE["value"],
                       "height": H, "width": W, "block_size": 128, "active_blocks": num_exec, "total_blocks": G,
                       "clip_length": L, "streams_per_gpu": S, "batch": B, "policy": args.policy,
                       "cuda_graphs": not args.no_graphs,
                       "l2_note": "kernel microbenchmarks rotate 8 plane sets (268 MB > 126 MB L2); the frame loop "
                                  "cycles 30 distinct 12.6 MB frames over 0.27 GB of planes per stream"},
            "e2e": e2e, "e2e_fp16": e2e_fp16, "gpu_launches": int(launches), "batched": batched, "config4": config4,
            "rl_semseg": rl, "cpu_baseline": cpu, "reference_gpu_path": ref_gpu, "reference_gpu_path_rl_semseg": ref_gpu_rl,
            "clocks": clk.summary(),
        }
        if kern is not None:
            line.update(roofline_records(kern, peaks))
            line["kernels"] = kern
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


def roofline_records(kern, peaks):
    """`roofline`: the kernel with the largest share of the step (conv_igemm_persistent_kernel<128,5>: 16 of the 43
    launches, ~44 % of the frame's GPU time, profiles/r02_frame_launch_shares.md) over its three characteristic
    shapes (layer2 / layer3 / layer4 3x3 convs at E = 40: 3.02 GFLOP each), sum of flops / sum of launch times.
    `roofline_l20` (largest single launch) and `roofline_hbm` (block movement) beside it."""
    dom_keys = ("conv3x3_c128_bs16(layer2)", "conv3x3_c256_bs8(layer3)", "conv3x3_c512_bs4(layer4)")
    flops = sum(kern[k]["flops"] for k in dom_keys)
    us = sum(kern[k]["us"] for k in dom_keys)
    tf = flops / us * 1e-6
    l20 = kern["conv3x3_c128_bs32(#20)"]
    hbm = kern["gather_halo_nhwc"]
    return {
        "roofline": {"bound": "tensor",
                     "kernel": "bc_conv_igemm / conv_igemm_persistent_kernel<128,5> (share-dominant kernel of the step) on the "
                               "3x3 convs of layer2 (128ch, 16-px blocks), layer3 (256ch, 8-px, split-K) and layer4 "
                               "(512ch, 4-px, split-K), E=40",
                     "achieved": tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": tf / peaks["tf_burst"],
                     "traffic": 3.42e6, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of the layer2 "
                     "shape (algorithmic: 6.2 MB; inputs come from L2), ncu --set full: profiles/r02b_conv_persist_ncu.md",
                     "per_shape": {k: kern[k] for k in dom_keys},
                     "same_shapes_8_streams_batched": (lambda b: {
                         "frac": sum(b[k]["flops"] for k in dom_keys) / sum(b[k]["us"] for k in dom_keys) * 1e-6 / peaks["tf_burst"],
                         "per_shape": {k: b[k] for k in dom_keys}})(kern["at_8_streams_batched"]) if "at_8_streams_batched" in kern else None,
                     "peak_source": peaks["source"] + " (burst: kernel timed alone)",
                     "algorithmic_flops_per_launch": flops / len(dom_keys), "us_per_launch": us / len(dom_keys)},
        "roofline_l20": {"bound": "tensor", "kernel": "bc_conv_igemm 3x3 128->128 on 32-px blocks, E=40 (SwiftNet layer #20: "
                                                      "the largest single launch of the step)",
                         "achieved": l20["tflops"], "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                         "frac": l20["tflops"] / peaks["tf_burst"], "algorithmic_flops_per_launch": l20["flops"],
                         "us_per_launch": l20["us"]},
        "roofline_hbm": {"bound": "hbm", "kernel": "bc_gather_halo TMA path (NHWC, C=128, BS=32, p=1, E=38: BASELINE config 2)",
                         "achieved": hbm["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm["gbs"] / peaks["hbm_gbs"],
                         "peak_source": peaks["source"], "algorithmic_bytes_per_launch": hbm["bytes"], "us_per_launch": hbm["us"],
                         "nchw": kern.get("gather_halo_nchw"),
                         "at_config5_size": {k: kern[k] for k in ("gather_halo_nhwc_cfg5", "gather_nhwc_cfg5",
                                                                  "scatter_nhwc_cfg5") if k in kern}},
    }


# =============================================================================================== microbench
def microbench(device, peaks, reps=200, sets=8):
    """BASELINE config 2 (1x128x256x512 fp16, 32-px feature blocks, E=38 of 128): achieved GB/s of
    every block kernel.  `sets` distinct plane/tile sets are rotated (8 x 33.5 MB planes > L2) and
    `reps` launches are timed between two CUDA events on the launching stream."""
    from blockcopy import _C

    N, C, H, W, BS, E, pad = 1, 128, 256, 512, 32, 38, 1
    g = torch.Generator(device=device).manual_seed(0)
    cells = torch.randperm(128, generator=torch.Generator().manual_seed(0))[:E]
    grid = torch.zeros(128, dtype=torch.bool)
    grid[cells] = True
    grid = grid.view(1, 1, 8, 16).to(device)
    gi = torch.empty(grid.shape, dtype=torch.int32, device=device)
    me = torch.empty(128, dtype=torch.int32, device=device)
    cnt = torch.empty(2, dtype=torch.int32, device=device)
    _C.compact_mask(grid.view(torch.uint8), gi, me, cnt)
    me = me[:E]
    res = {}
    tile_b = C * BS * BS * 2
    algo = {"gather": 2 * E * tile_b, "gather_halo": 2 * E * C * (BS + 2 * pad) ** 2 * 2, "scatter": 2 * E * tile_b,
            "copy_blocks": 2 * C * H * W * 2}
    for lay, fmt in (("nhwc", torch.channels_last), ("nchw", torch.contiguous_format)):
        planes = [torch.randn(N, C, H, W, device=device, dtype=torch.float16, generator=g).contiguous(memory_format=fmt)
                  for _ in range(sets)]
        outs = [torch.empty_like(planes[0]) for _ in range(2)]
        tiles = [torch.randn(E, C, BS, BS, device=device, dtype=torch.float16, generator=g).contiguous(memory_format=fmt)
                 for _ in range(sets)]
        padded = [torch.empty(E, C, BS + 2 * pad, BS + 2 * pad, device=device, dtype=torch.float16).contiguous(memory_format=fmt)
                  for _ in range(sets)]
        ops = {
            "gather": lambda i: _C.gather(tiles[i % sets], planes[i % sets], me, E),
            "gather_halo": lambda i: _C.gather_halo(padded[i % sets], planes[i % sets], me, E, BS, pad),
            "scatter": lambda i: _C.scatter(tiles[i % sets], planes[i % sets], me, E),
            "copy_blocks": lambda i: _C.copy_blocks(outs[i % 2], planes[i % sets], tiles[i % sets], gi),
        }
        for tma in (True, False):
            _C.set_tma_enabled(tma)
            for name, fn in ops.items():
                if name == "copy_blocks" and not tma:
                    continue
                for i in range(sets):
                    fn(i)
                torch.cuda.synchronize()
                # one CUDA graph of `reps` back-to-back launches: host launch cost is not what is measured
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for i in range(reps):
                        fn(i)
                graph.replay()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                graph.replay()
                b.record()
                torch.cuda.synchronize()
                us = a.elapsed_time(b) * 1e3 / reps
                key = f"{name}_{lay}" + ("" if tma else "_simt")
                res[key] = {"us": us, "bytes": algo[name], "gbs": algo[name] / us * 1e-3,
                            "frac_of_hbm_peak": algo[name] / us * 1e-3 / peaks["hbm_gbs"]}
                del graph
        _C.set_tma_enabled(True)
        del planes, tiles, padded, outs
    res.update(large_gather_microbench(device, peaks))
    res.update(conv_microbench(device, peaks, me_full=None))
    # the same conv launches when 8 streams are batched along N (config 4: E = 8 x 40 blocks): how far the kernels get
    # when the grid fills the GPU
    res["at_8_streams_batched"] = conv_microbench(device, peaks, me_full=None, reps=20, sets=2, E=320, images=8)
    res.update(io_microbench(device, peaks))
    return res


def io_microbench(device, peaks, reps=100, sets=8):
    """Driver-side steps at BASELINE size: uint8 1024x2048 frame -> fp16 input
    (18.9 MB algorithmic: 6.3 read + 12.6 written), and (1,19,256,512) fp16 logits -> 1024x2048 uint8 label map
    (7.1 MB: 5.0 read + 2.1 written), next to the torch op sequences of the reference's driver.  Rotating
    `sets` buffer sets (151 / 97 MB), graph-timed."""
    from consumers.frame_io import FrameNormalizer, predict_labels
    import torch.nn.functional as F

    g = torch.Generator(device=device).manual_seed(0)
    H, W = 1024, 2048
    u8 = [torch.randint(0, 256, (1, H, W, 3), dtype=torch.uint8, device=device, generator=g) for _ in range(sets)]
    fin = [torch.empty(1, 3, H, W, dtype=torch.float16, device=device) for _ in range(sets)]
    logits = [torch.randn(1, 19, H // 4, W // 4, device=device, generator=g).half() for _ in range(sets)]
    lab = [torch.empty(1, H, W, dtype=torch.uint8, device=device) for _ in range(sets)]
    norm = FrameNormalizer()
    mean = torch.as_tensor(norm.mean, dtype=torch.float32, device=device).view(1, 3, 1, 1)
    std = torch.as_tensor(norm.std, dtype=torch.float32, device=device).view(1, 3, 1, 1)
    ops = {
        "frame_from_u8": (lambda i: norm(u8[i % sets], out=fin[i % sets]), H * W * 3 + H * W * 3 * 2),
        "frame_from_u8_torch_ops": (lambda i: u8[i % sets].permute(0, 3, 1, 2).float().div_(255).sub_(mean).div_(std).half(),
                                    H * W * 3 + H * W * 3 * 2),
        "upsample4x_argmax": (lambda i: predict_labels(logits[i % sets], out=lab[i % sets]), 19 * H * W // 16 * 2 + H * W),
        "upsample4x_argmax_torch_ops": (lambda i: F.interpolate(logits[i % sets], size=(H, W), mode="bilinear").max(dim=1)[1],
                                        19 * H * W // 16 * 2 + H * W),
    }
    res = {}
    for name, (fn, nbytes) in ops.items():
        r = reps if not name.endswith("torch_ops") else 20
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(r):
                fn(i)
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / r
        res[name] = {"us": us, "bytes": nbytes, "gbs": nbytes / us * 1e-3, "frac_of_hbm_peak": nbytes / us * 1e-3 / peaks["hbm_gbs"]}
        del graph
    return res


def large_gather_microbench(device, peaks, reps=100, sets=2):
    """Same kernels at BASELINE config-5 size (2048x4096 image => 128x512x1024 plane = 134 MB, 154 of 512 blocks,
    ~90 MB per launch): long enough that the launch ramp no longer dominates."""
    from blockcopy import _C

    C, H, W, BS, E, G = 128, 512, 1024, 32, 154, 512
    cells = torch.randperm(G, generator=torch.Generator().manual_seed(0))[:E].sort().values.to(torch.int32).to(device)
    fmt = torch.channels_last
    planes = [torch.randn(1, C, H, W, device=device, dtype=torch.float16).contiguous(memory_format=fmt) for _ in range(sets)]
    padded = [torch.empty(E, C, BS + 2, BS + 2, device=device, dtype=torch.float16).contiguous(memory_format=fmt) for _ in range(sets)]
    tiles = [torch.empty(E, C, BS, BS, device=device, dtype=torch.float16).contiguous(memory_format=fmt) for _ in range(sets)]
    out = {}
    for name, fn, nbytes in (
            ("gather_halo_nhwc_cfg5", lambda i: _C.gather_halo(padded[i % sets], planes[i % sets], cells, E, BS, 1), 2 * E * C * (BS + 2) ** 2 * 2),
            ("gather_nhwc_cfg5", lambda i: _C.gather(tiles[i % sets], planes[i % sets], cells, E), 2 * E * C * BS * BS * 2),
            ("scatter_nhwc_cfg5", lambda i: _C.scatter(tiles[i % sets], planes[i % sets], cells, E), 2 * E * C * BS * BS * 2)):
        for i in range(sets):
            fn(i)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(reps):
                fn(i)
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps
        out[name] = {"us": us, "bytes": nbytes, "gbs": nbytes / us * 1e-3, "frac_of_hbm_peak": nbytes / us * 1e-3 / peaks["hbm_gbs"]}
        del graph
    return out


def conv_microbench(device, peaks, me_full=None, reps=50, sets=4, E=40, images=1):
    """bc_conv_igemm on the four characteristic 3x3 layers of SwiftNet-RN18 at 1024x2048 (grid 8x16,
    E = 40 executed blocks): achieved TFLOP/s = 2*9*Cin*Cout*BS^2*E / time, graph-timed, rotating
    `sets` plane copies (the layer-#20 planes are 33.5 MB each)."""
    from blockcopy import _C

    out = {}
    cells = torch.randperm(128 * images, generator=torch.Generator().manual_seed(0))[:E].sort().values.to(torch.int32).to(device)
    g = torch.Generator(device=device).manual_seed(1)
    for name, Cin, Cout, BS in (("conv3x3_c128_bs32(#20)", 128, 128, 32), ("conv3x3_c64_bs32(layer1)", 64, 64, 32),
                                ("conv3x3_c128_bs16(layer2)", 128, 128, 16), ("conv3x3_c256_bs8(layer3)", 256, 256, 8),
                                ("conv3x3_c512_bs4(layer4)", 512, 512, 4)):
        H, W = 8 * BS, 16 * BS
        planes = [torch.randn(images, Cin, H, W, device=device, dtype=torch.float16, generator=g).contiguous(memory_format=torch.channels_last)
                  for _ in range(sets)]
        w = (torch.randn(Cout, Cin, 3, 3, device=device, dtype=torch.float16, generator=g) * 0.05).contiguous(memory_format=torch.channels_last)
        bias = torch.zeros(Cout, device=device, dtype=torch.float16)
        outs = [torch.empty(E, Cout, BS, BS, device=device, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
                for _ in range(2)]
        fn = lambda i: _C.conv_igemm(outs[i % 2], planes[i % sets], w, bias, None, cells, E, BS, 1, 1)  # noqa: E731
        for i in range(sets):
            fn(i)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(reps):
                fn(i)
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps
        flops = 2.0 * 9 * Cin * Cout * BS * BS * E
        out[name] = {"us": us, "flops": flops, "tflops": flops / us * 1e-6,
                     "frac_of_tensor_peak": flops / us * 1e-6 / peaks["tf_burst"]}
        del graph, planes, outs
    return out


# =============================================================================================== CPU baseline
def cpu_dense_baseline(H, W, frames=5, warm=2):
    """BASELINE config 1: dense SwiftNet-RN18 forward, fp32, CPU torch on all host cores (the only
    part of the reference that runs without CUDA).  Uses the reference's own model code when it is
    staged in baseline/_ref (kind 'reference'), else this repo's architecture-identical consumer
    (kind 'port')."""
    kind, net = "port", None
    ref_ss = os.path.join(ROOT, "baseline", "_ref", "semantic_segmentation")
    if os.path.isdir(os.path.join(ref_ss, "lib", "models", "swiftnet")):
        try:  # the reference's own model code (staged copy); `blockcopy` it imports is only used for a decorator
            import blockcopy  # noqa: F401
            sys.path.insert(0, ref_ss)
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                from lib.models.swiftnet.backbones.resnet import resnet18
                from lib.models.swiftnet.swiftnet import SwiftNet
                from lib.utils import bn_fusion
                torch.manual_seed(0)
                net = SwiftNet(resnet18(pretrained=False), num_classes=19, num_features=128, use_spp=True).eval()
                net = bn_fusion.fuse_bn_recursively(net)
            kind = "reference"
        except Exception:
            net = None
    if net is None:
        from consumers.swiftnet_rn18 import build_swiftnet_rn18

        net = build_swiftnet_rn18(seed=0)
    x = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(0))
    if "cpus" in _ORIG_AFFINITY:  # the CPU baseline uses ALL host cores, not only the GPU's NUMA node
        try:
            os.sched_setaffinity(0, _ORIG_AFFINITY["cpus"])
        except Exception:
            pass
    try:  # torchrun exports OMP_NUM_THREADS=1: use every core this process may run on
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    cores = torch.get_num_threads()
    times = []
    with torch.no_grad():
        for i in range(warm + frames):
            t0 = time.perf_counter()
            net(x)
            if i >= warm:
                times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": 1.0 / med, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"dense SwiftNet-RN18 fp32 forward, 1x3x{H}x{W}, {warm} warm-up + {frames} timed, median"}


def bench_reference(args):
    """Reference arm: the reference's CPU-runnable path (dense SwiftNet-RN18, CPU torch, all cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = max(3, min(args.steps, 20))
    warm = max(1, min(args.warmup, 2))
    cpu = cpu_dense_baseline(args.height, args.width, frames=frames, warm=warm)
    line = {"impl": "reference", "metric": "frames/s @1024x2048, 30% active blocks (SwiftNet-RN18 + BlockCopy)",
            "value": cpu["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": frames, "warmup": warm,
            "ms_per_step": 1000.0 / cpu["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[0]: dense SwiftNet-RN18 forward on CPU torch (the reference's block "
                                   "path is CUDA-only)", "height": args.height, "width": args.width},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        return bench_reference(args)
    if args.microbench:
        assert torch.cuda.is_available()
        print(json.dumps({"kernels": microbench(torch.device("cuda", 0), load_peaks())}))
        return
    bench_ours(args)


if __name__ == "__main__":
    main()
