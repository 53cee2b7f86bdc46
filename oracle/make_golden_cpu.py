"""oracle/make_golden_cpu.py -- TEST INFRASTRUCTURE.  Run in the build container (needs
/root/reference):   python oracle/make_golden_cpu.py

Imports the UNMODIFIED reference (cupy stubbed, its four CUDA kernels replaced by the C oracle,
everything else -- wrapper, FIFO protocol, policy, SwiftNet -- is the reference's own Python on
CPU torch) and writes the fixtures the test-suite compares against:

  tests/golden/index_kat.json        grid_idx / mapping_exec / transfer_idx of seeded mask
                                     sequences, from the reference's get_grid_mappings and
                                     BlockFeatures._process_grid (tensorwrapper.py:108-178)
  tests/golden/swiftnet_cpu_clip.pt  reference SwiftNet-RN18 + BlockCopyModel, fp32, 256x512,
                                     64-px blocks (grid 4x8), 6 frames, replayed seeded masks:
                                     per-frame argmax, strided logits, full logits of 2 frames
  tests/golden/policy_cpu.pt         PolicyNet features / logits and InformationGainSemSeg on
                                     seeded inputs (weights by deterministic_init_)
"""
import json
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_env  # noqa: E402

ref = ref_env.import_reference("cpu")
sys.path.append(os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200"))  # consumers/ only
from consumers.clips import PolicyReplay, deterministic_init_, synthetic_clip  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
torch.set_num_threads(8)


def settings(**kw):
    s = dict(block_policy="all", block_num_classes=19, block_optim_lr=1e-4, block_optim_wd=1e-3,
             block_optim_momentum=0, block_target=0.5, block_complexity_weight=5, block_size=128,
             block_train_interval=4, block_cost_momentum=0.9, block_policy_verbose=False)
    s.update(kw)
    return s


# ------------------------------------------------------------------------------------------- 1. index KATs
def index_kats():
    from blockcopy.core.tensorwrapper import BlockFeatures, get_grid_mappings

    cases = []
    g = torch.Generator().manual_seed(0)
    seqs = {
        "tiny": [torch.ones(1, 1, 2, 4, dtype=torch.bool),
                 torch.tensor([[1, 0, 1, 0], [0, 1, 1, 0]], dtype=torch.bool).view(1, 1, 2, 4),
                 torch.tensor([[0, 0, 1, 1], [1, 0, 0, 0]], dtype=torch.bool).view(1, 1, 2, 4)],
    }
    for name, shape in (("n1_8x16", (1, 1, 8, 16)), ("n2_4x8", (2, 1, 4, 8)), ("n3_3x5", (3, 1, 3, 5))):
        seq = [torch.ones(shape, dtype=torch.bool)]
        for frac in (0.3, 0.5, 0.05, 1.0, 0.0, 0.7):
            seq.append(torch.rand(shape, generator=g) < frac)
        seqs[name] = seq
    for name, seq in seqs.items():
        prev = None
        frames = []
        for grid in seq:
            n_exec = int(grid.sum())
            gi, me = get_grid_mappings(n_exec, grid, ~grid, grid.numel(), grid.numel() - n_exec)
            bf = BlockFeatures(device="cpu")
            bf._process_grid(grid, prev)
            assert torch.equal(bf._grid_idx, gi) and torch.equal(bf._mapping_exec, me.int())
            frames.append(dict(grid=grid.int().flatten().tolist(), shape=list(grid.shape),
                               grid_idx=bf._grid_idx.flatten().tolist(),
                               mapping_exec=bf._mapping_exec.tolist(),
                               transfer_idx=None if prev is None else bf._transfer_idx.tolist()))
            prev = bf
        cases.append(dict(name=name, frames=frames))
    with open(os.path.join(GOLD, "index_kat.json"), "w") as f:
        json.dump(cases, f)
    print("index_kat.json:", [(c["name"], len(c["frames"])) for c in cases])


# ------------------------------------------------------------------------------------------- 2. SwiftNet clip
def swiftnet_clip():
    from lib.models.swiftnet.backbones.resnet import resnet18
    from lib.models.swiftnet.swiftnet import SwiftNet
    from lib.utils import bn_fusion

    H, W, BS, T = 256, 512, 64, 6
    torch.manual_seed(0)
    random.seed(0)
    net = SwiftNet(resnet18(pretrained=False), num_classes=19, num_features=128, use_spp=True).eval()
    deterministic_init_(net, seed=0)
    model = ref.BlockCopyModel(net, settings(block_size=BS)).eval()
    model = bn_fusion.fuse_bn_recursively(model)
    # seeded masks: frame 0 all, then a varying number of blocks incl. one empty and one full frame
    g = torch.Generator().manual_seed(1)
    grids = [torch.ones(1, 1, H // BS, W // BS, dtype=torch.bool)]
    for frac in (0.3, 0.5, 0.0, 0.25, 1.0):
        cells = grids[0].numel()
        m = torch.zeros(cells, dtype=torch.bool)
        m[torch.randperm(cells, generator=g)[: round(frac * cells)]] = True
        grids.append(m.view_as(grids[0]))
    model.policy = PolicyReplay(BS, grids)
    clip = synthetic_clip(T, H, W, seed=3, dtype=torch.float32)
    outs, states = [], []
    with torch.no_grad():
        model.reset_temporal()
        for t in range(T):
            out = model(clip[t])
            outs.append(out.clone())
            states.append(model.policy_meta["frame_state"].clone())
    fix = dict(H=H, W=W, BS=BS, T=T, clip_seed=3, init_seed=0,
               grids=torch.stack(grids).to(torch.uint8),
               argmax=torch.stack([o.argmax(1).to(torch.uint8) for o in outs]),
               logits_strided=torch.stack([o[:, :, ::4, ::4].clone() for o in outs]),
               logits_full={t: outs[t].clone() for t in (0, T - 2)},
               logits_abs_mean=[float(o.abs().mean()) for o in outs],
               frame_state_sum=[float(s.double().sum()) for s in states])
    torch.save(fix, os.path.join(GOLD, "swiftnet_cpu_clip.pt"))
    print("swiftnet_cpu_clip.pt: |logits| mean per frame", [round(v, 4) for v in fix["logits_abs_mean"]],
          "range", float(outs[-1].min()), float(outs[-1].max()))


# ------------------------------------------------------------------------------------------- 3. policy pieces
def policy_inputs(seed=5, N=1, H=256, W=512, BS=64, K=19):
    """Seeded inputs of the policy fixtures; regenerated (not stored) by tests/test_policy.py."""
    g = torch.Generator().manual_seed(seed)
    meta = dict(inputs=torch.randn(N, 3, H, W, generator=g),
                frame_state=torch.randn(N, 3, H, W, generator=g),
                output_repr=torch.randn(N, K, H // 4, W // 4, generator=g),
                grid=torch.rand(N, 1, H // BS, W // BS, generator=g) < 0.4)
    ig_meta = dict(outputs=torch.randn(N, K, H // 4, W // 4, generator=g),
                   outputs_prev=torch.randn(N, K, H // 4, W // 4, generator=g))
    return meta, ig_meta


def policy_pieces():
    from blockcopy.policy.information_gain import InformationGainSemSeg
    from blockcopy.policy.net import PolicyNet

    BS, K = 64, 19
    meta, ig_meta = policy_inputs(BS=BS, K=K)
    net = PolicyNet(block_size=BS, task_num_classes=K)
    deterministic_init_(net, seed=2)
    net.train()
    with torch.no_grad():
        logits = net(meta)
    ig = InformationGainSemSeg(K)(ig_meta)
    torch.save(dict(logits=logits, ig=ig, BS=BS, K=K, init_seed=2, input_seed=5),
               os.path.join(GOLD, "policy_cpu.pt"))
    print("policy_cpu.pt: logits", tuple(logits.shape), "ig", tuple(ig.shape))


# ------------------------------------------------------------------------------------------- 4. CSP stand-in
def csp_standin_clip(wide=False):
    """Second consumer (Pedestron CSPBlockCopy op set, tests/csp_standin.py) on the reference package."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from csp_standin import StandinDetector

    H, W, BS = 128, 256, 64
    det = StandinDetector(settings(block_size=BS), wide=wide).eval()
    deterministic_init_(det, seed=4)
    g = torch.Generator().manual_seed(2)
    grids = [torch.ones(1, 1, H // BS, W // BS, dtype=torch.bool)]
    for frac in (0.5, 0.25, 0.0, 0.75):
        cells = grids[0].numel()
        m = torch.zeros(cells, dtype=torch.bool)
        m[torch.randperm(cells, generator=g)[: round(frac * cells)]] = True
        grids.append(m.view_as(grids[0]))
    det.policy = PolicyReplay(BS, grids)
    clip = synthetic_clip(len(grids), H, W, seed=6, dtype=torch.float32)
    with torch.no_grad():
        outs = [det.simple_test(f).clone() for f in clip]
    torch.save(dict(H=H, W=W, BS=BS, init_seed=4, clip_seed=6, grids=torch.stack(grids).to(torch.uint8),
                    outs=torch.stack(outs)), os.path.join(GOLD, "csp_standin_wide_cpu.pt" if wide else "csp_standin_cpu.pt"))
    print("csp_standin_wide_cpu.pt:" if wide else "csp_standin_cpu.pt:", tuple(outs[0].shape), [round(float(o.abs().mean()), 4) for o in outs])


if __name__ == "__main__":
    csp_standin_clip()
    csp_standin_clip(wide=True)
    index_kats()
    policy_pieces()
    swiftnet_clip()
