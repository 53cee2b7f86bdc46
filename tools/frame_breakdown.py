"""Where does a frame go?  (a) GPU time of one replay of the E=40 graph, (b) wall time of the full
model() call (policy + copies + replay), (c) host-only time of model() with the GPU idle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))
import torch
import blockcopy
from blockcopy.core.argparser import default_settings
from consumers.clips import PolicyFixedFraction, synthetic_clip
from consumers.swiftnet_rn18 import build_swiftnet_rn18

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = blockcopy.BlockCopyModel(build_swiftnet_rn18(), default_settings(block_policy="all", block_cuda_graphs=True)).eval().to(dev).half()
model.policy = PolicyFixedFraction(128, 0.3, 8, 0)
clip = synthetic_clip(30, 1024, 2048, seed=0, device=dev)
with torch.no_grad():
    for rep in range(3):
        model.reset_temporal()
        for f in clip:
            model(f)
    torch.cuda.synchronize()
    for (E, _patch), entry in model._graphs.graphs.items():
        if E != model.policy.num_exec_for(128):
            continue
        g = entry[0]
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            g.replay()
        b.record(); torch.cuda.synchronize()
        print(f"graph E={E}: {a.elapsed_time(b) / 20 * 1000:.1f} us GPU per replay, kernels of ours {entry[3]}")
    # full call, steady frames
    model.reset_temporal(); model(clip[0]); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in clip[1:]:
        model(f)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"model() steady frames: host issue {1e6 * (t1 - t0) / 29:.1f} us/frame, wall incl. drain {1e6 * (t2 - t0) / 29:.1f} us/frame")
    # host-side pieces
    t0 = time.perf_counter()
    for _ in range(100):
        meta = model.policy({"inputs": clip[1], "outputs": 1, "outputs_prev": 1})
    torch.cuda.synchronize()
    print(f"policy.forward (host mask + H2D): {1e4 * (time.perf_counter() - t0):.1f} us")
