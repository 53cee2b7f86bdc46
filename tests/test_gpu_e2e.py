"""End-to-end parity on the GPU: SwiftNet-RN18 behind BlockCopyModel, CUDA kernels through the C
ABI, against (a) the fixture the UNMODIFIED reference produced on CPU in fp32 and (b) when present,
the fixture it produced on a B200 in fp16 through the NVRTC cupy shim."""
import glob
import os
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def _settings(**kw):
    from blockcopy.core.argparser import default_settings

    return default_settings(**kw)


def _loaded_native():
    with open("/proc/self/maps") as f:
        return "libblockcopy_sm100.so" in f.read()


@pytest.mark.parametrize("channels_last", [True, False])
def test_swiftnet_clip_fp32_vs_reference_cpu_fixture(golden_dir, channels_last):
    import blockcopy
    from consumers.clips import PolicyReplay, deterministic_init_, synthetic_clip
    from consumers.swiftnet_rn18 import SwiftNetRN18, fuse_conv_bn_

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fix = torch.load(os.path.join(golden_dir, "swiftnet_cpu_clip.pt"))
    H, W, BS, T = fix["H"], fix["W"], fix["BS"], fix["T"]
    net = deterministic_init_(SwiftNetRN18().eval(), seed=fix["init_seed"])
    model = blockcopy.BlockCopyModel(net, _settings(block_policy="all", block_size=BS,
                                                    block_channels_last=channels_last)).eval()
    fuse_conv_bn_(model)
    model = model.cuda()
    model.policy = PolicyReplay(BS, list(fix["grids"].bool()))
    clip = synthetic_clip(T, H, W, seed=fix["clip_seed"], dtype=torch.float32, device="cuda")
    with torch.no_grad():
        model.reset_temporal()
        for t in range(T):
            out = model(clip[t])
            assert out.is_cuda and tuple(out.shape) == (1, 19, H // 4, W // 4)
            scale = fix["logits_abs_mean"][t]
            # fp32 storage + fp32 accumulate on both sides: only summation order differs
            err = (out[:, :, ::4, ::4].cpu() - fix["logits_strided"][t]).abs().max().item()
            assert err <= 1e-3 * scale, (t, err, scale)
            agree = (out.argmax(1).to(torch.uint8).cpu() == fix["argmax"][t]).float().mean().item()
            assert agree >= 0.999, (t, agree)
            fs = model.policy_meta["frame_state"]
            assert abs(float(fs.double().sum()) - fix["frame_state_sum"][t]) < 1e-3  # bit-exact block movement
    assert len(model.block_temporal_features._planes) == 21
    assert _loaded_native(), "native library not loaded: the CUDA path did not run"


def test_swiftnet_clip_fp16_vs_reference_gpu_fixture(golden_dir):
    """Tolerance (stated): fp16 storage, fp32 accumulate => |d| <= 2^-8 * max|ref| + 1e-3 per
    logit, identical argmax on >= 99.9 % of pixels (BASELINE.json north star)."""
    import blockcopy
    from consumers.clips import PolicyReplay, deterministic_init_, synthetic_clip
    from consumers.swiftnet_rn18 import SwiftNetRN18, fuse_conv_bn_

    path = os.path.join(golden_dir, "swiftnet_gpu_fp16_clip.pt")
    if not os.path.exists(path):
        pytest.skip("swiftnet_gpu_fp16_clip.pt not generated yet (oracle/make_golden_gpu.py)")
    fix = torch.load(path)
    H, W, BS, T = fix["H"], fix["W"], fix["BS"], fix["T"]
    net = deterministic_init_(SwiftNetRN18().eval(), seed=fix["init_seed"], gain=fix.get("init_gain", 1.0))
    model = blockcopy.BlockCopyModel(net, _settings(block_policy="all", block_size=BS)).eval()
    fuse_conv_bn_(model)
    model = model.cuda().half()
    model.policy = PolicyReplay(BS, list(fix["grids"].bool()))
    clip = synthetic_clip(T, H, W, seed=fix["clip_seed"], dtype=torch.float16, device="cuda")
    with torch.no_grad():
        model.reset_temporal()
        for t in range(T):
            out = model(clip[t]).float().cpu()
            ref = fix["logits_strided"][t].float()
            tol = 2 ** -8 * float(ref.abs().max()) + 1e-3
            err = (out[:, :, ::4, ::4] - ref).abs().max().item()
            assert err <= tol, (t, err, tol)
            agree = (out.argmax(1).to(torch.uint8) == fix["argmax"][t]).float().mean().item()
            assert agree >= 0.999, (t, agree)
            fs = model.policy_meta["frame_state"]
            assert torch.equal(fs[:, :, ::8, ::8].cpu(), fix["frame_state_strided"][t]), "block movement is bit-exact"


def test_benchmarked_config_graph_mode_vs_reference_gpu_fixture(golden_dir):
    """The configuration bench.py times -- 1024x2048, 128-px blocks, 30 frames, frame 0 all blocks then 40 of 128,
    `block_cuda_graphs=True` (one-tile 320-CTA conv form, persistent 80/120-CTA forms, split-K clusters with the
    model-owned scratch, side-stream graph branches, all together) -- against what the UNMODIFIED reference produced
    for the same clip and masks on a B200 (oracle/make_golden_gpu.py full -> swiftnet_gpu_fp16_full.npz).
    Stated tolerance: fp16 storage, fp32 accumulate => |d| <= 2^-8 * max|ref| + 1e-3 per logit; argmax equal on
    >= 99.9 % of the pixels of every frame; frame_state (block movement) bit-exact.  And graph replays == eager
    launches bit for bit at this size."""
    import numpy as np

    import blockcopy
    from consumers.clips import PolicyFixedFraction, deterministic_init_, synthetic_clip
    from consumers.swiftnet_rn18 import SwiftNetRN18, fuse_conv_bn_

    path = os.path.join(golden_dir, "swiftnet_gpu_fp16_full.npz")
    if not os.path.exists(path):
        pytest.skip("swiftnet_gpu_fp16_full.npz not generated yet (oracle/make_golden_gpu.py full)")
    fix = np.load(path)
    H, W, BS, T = int(fix["H"]), int(fix["W"]), int(fix["BS"]), int(fix["T"])
    assert (H, W, BS, T) == (1024, 2048, 128, 30)
    clip = synthetic_clip(T, H, W, seed=int(fix["clip_seed"]), dtype=torch.float16, device="cuda")

    def build(graphs):
        net = deterministic_init_(SwiftNetRN18().eval(), seed=int(fix["init_seed"]), gain=float(fix["init_gain"]))
        model = blockcopy.BlockCopyModel(net, _settings(block_policy="all", block_size=BS, block_cuda_graphs=graphs)).eval()
        fuse_conv_bn_(model)
        model = model.cuda().half()
        model.policy = PolicyFixedFraction(BS, fraction=0.3, quantize=8, seed=int(fix["mask_seed"]))
        return model

    def run(model):
        outs = []
        model.policy.reseed(int(fix["mask_seed"]))
        with torch.no_grad():
            model.reset_temporal()
            for t in range(T):
                o = model(clip[t])
                g = model.policy_meta["grid"].cpu().numpy().astype(np.uint8)
                assert (g == fix["grids"][t]).all(), f"frame {t}: mask differs from the fixture's"
                outs.append((o.clone(), model.policy_meta["frame_state"].clone()))
        torch.cuda.synchronize()
        return outs

    graphed = build(True)
    run(graphed)          # eager pass: planes are allocated, each block count is seen once
    run(graphed)          # capture pass
    replayed = run(graphed)
    assert {k[0] for k in graphed._graphs.graphs} == {128, 40}, "graphs were not captured"
    worst_err, worst_agree = 0.0, 1.0
    for t, (o, fs) in enumerate(replayed):
        ref = torch.from_numpy(fix["logits_strided"][t]).float()
        got = o[0, :, (t % 8)::8, ((3 * t) % 8)::8].float().cpu()
        tol = 2 ** -8 * float(fix["logits_abs_max"][t]) + 1e-3
        err = (got - ref).abs().max().item()
        assert err <= tol, (t, err, tol)
        agree = (o.argmax(1)[0].to(torch.uint8).cpu() == torch.from_numpy(fix["argmax"][t])).float().mean().item()
        assert agree >= 0.999, (t, agree)
        want = torch.from_numpy(fix["frame_state_strided"][t])
        assert torch.equal(fs[0, :, (t % 16)::16, ((5 * t) % 16)::16].cpu(), want), f"frame {t}: block movement differs"
        worst_err, worst_agree = max(worst_err, err / tol), min(worst_agree, agree)
    print(f"benchmarked config vs reference fixture: worst err/tol {worst_err:.3f}, worst argmax agreement {worst_agree:.5f}")
    eager = run(build(False))
    for t, ((a, fa), (b, fb)) in enumerate(zip(eager, replayed)):
        assert torch.equal(a, b), (t, float((a.float() - b.float()).abs().max()))
        assert torch.equal(fa, fb), t
    assert _loaded_native()


def test_rl_semseg_runs_and_trains_fp16():
    """The reference driver's configuration: model in fp16, policy net in fp32 (test_swiftnet.py:118-123)."""
    import blockcopy
    from consumers.clips import synthetic_clip
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    random.seed(0)
    torch.manual_seed(0)
    model = blockcopy.BlockCopyModel(build_swiftnet_rn18(), _settings(block_policy="rl_semseg", block_target=0.3,
                                                                    block_train_interval=3)).eval().cuda().half()
    model.policy.net = model.policy.net.float().train()
    before = [p.detach().clone() for p in model.policy.net.parameters()]
    clip = synthetic_clip(7, 512, 1024, seed=2, dtype=torch.float16, device="cuda")
    with torch.no_grad():
        model.reset_temporal()
        for f in clip:
            out = model(f)
    assert out.dtype == torch.float16 and tuple(out.shape) == (1, 19, 128, 256) and torch.isfinite(out).all()
    assert any(not torch.equal(a, b) for a, b in zip(before, model.policy.net.parameters()))
    assert model.policy_meta["num_exec"] % 2 == 0  # 32 blocks / 16


def test_noblocks_and_reset_isolate_clips():
    """reset_temporal starts from a clean slate: two clips processed back to back give the same
    outputs as each processed alone (per-stream state lives in the wrapper, not in globals)."""
    import blockcopy
    from consumers.clips import PolicyFixedFraction, synthetic_clip
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    def run(model, clip, seed):
        model.policy.reseed(seed)
        model.reset_temporal()
        with torch.no_grad():
            return [model(f).clone() for f in clip]

    model = blockcopy.BlockCopyModel(build_swiftnet_rn18(), _settings(block_policy="all", block_size=64)).eval().cuda().half()
    model.policy = PolicyFixedFraction(64, fraction=0.3, quantize=2, seed=0)
    a = synthetic_clip(4, 256, 512, seed=1, device="cuda")
    b = synthetic_clip(4, 256, 512, seed=2, device="cuda")
    ra1, rb1 = run(model, a, 5), run(model, b, 6)
    rb2, ra2 = run(model, b, 6), run(model, a, 5)
    for x, y in zip(ra1 + rb1, ra2 + rb2):
        assert torch.equal(x, y)


def test_cuda_graph_mode_is_bit_identical_to_eager():
    """block_cuda_graphs=True: first occurrence of a block count runs eagerly, the second is
    captured, later ones are replayed -- outputs, frame_state and policy_meta semantics must equal
    the eager mode bit for bit, across a clip boundary (reset_temporal keeps the planes)."""
    import blockcopy
    from consumers.clips import PolicyReplay, synthetic_clip
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    BS, H, W = 64, 256, 512
    g = torch.Generator().manual_seed(0)
    grids = [torch.ones(1, 1, 4, 8, dtype=torch.bool)]
    for e in (8, 8, 12, 8, 0, 12, 8, 32, 8):
        m = torch.zeros(32, dtype=torch.bool)
        m[torch.randperm(32, generator=g)[:e]] = True
        grids.append(m.view(1, 1, 4, 8))
    clip = synthetic_clip(len(grids), H, W, seed=4, device="cuda")

    def run(graphs):
        model = blockcopy.BlockCopyModel(build_swiftnet_rn18(), _settings(block_policy="all", block_size=BS,
                                                                        block_cuda_graphs=graphs)).eval().cuda().half()
        model.policy = PolicyReplay(BS, grids)
        outs = []
        with torch.no_grad():
            for _ in range(2):  # two clips: the second one replays graphs captured during the first
                model.reset_temporal()
                model.policy.rewind()
                prev = None
                for f in clip:
                    o = model(f)
                    assert model.policy_meta["outputs"] is o
                    if prev is not None and model.policy_meta["num_exec"]:
                        assert model.policy_meta["outputs_prev"] is not o
                    outs.append((o.clone(), model.policy_meta["frame_state"].clone()))
                    prev = o
        return model, outs

    m_eager, eager = run(False)
    m_graph, graphed = run(True)
    assert len(m_graph._graphs.graphs) >= 3, "graphs were not captured"
    for t, ((a, fa), (b, fb)) in enumerate(zip(eager, graphed)):
        assert torch.equal(a, b), f"frame {t}: outputs differ between eager and graph mode"
        assert torch.equal(fa, fb), f"frame {t}: frame_state differs"


def test_concurrent_cuda_streams_equal_sequential():
    """Independent video streams on their own CUDA streams (bench.py --concurrent-streams): all state is per
    wrapper object (planes, graphs, and in graph mode the split-K scratch whose address the graphs bake in) or
    per CUDA stream (eager split-K scratch), so the outputs must equal the one-after-the-other run bit for bit."""
    import blockcopy
    from consumers.clips import PolicyFixedFraction, synthetic_clip
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    BS, H, W, S, T = 64, 256, 512, 3, 5
    clips = [synthetic_clip(T, H, W, seed=10 + s, device="cuda") for s in range(S)]

    def build():
        models = []
        for s in range(S):
            m = blockcopy.BlockCopyModel(build_swiftnet_rn18(seed=0), _settings(block_policy="all", block_size=BS,
                                                                              block_cuda_graphs=True)).eval().cuda().half()
            m.policy = PolicyFixedFraction(BS, fraction=0.3, quantize=2, seed=s)
            models.append(m)
        return models

    def run(models, streams):
        outs = [[] for _ in range(S)]
        with torch.no_grad():
            for rep in range(3):  # clip 0: eager, clip 1: capture, clip 2: replay
                for s, m in enumerate(models):
                    m.reset_temporal()
                    m.policy.reseed(s)
                for t in range(T):
                    for s, m in enumerate(models):
                        if streams is None:
                            o = m(clips[s][t])
                        else:
                            with torch.cuda.stream(streams[s]):
                                o = m(clips[s][t])
                        if rep == 2:
                            outs[s].append(o.clone() if streams is None else o)
        torch.cuda.synchronize()
        return [[o.clone() for o in per] for per in outs] if streams is not None else outs

    seq = run(build(), None)
    models = build()  # parameters are uploaded / converted on the current stream ...
    side = [torch.cuda.Stream() for _ in range(S)]
    for st in side:
        st.wait_stream(torch.cuda.current_stream())  # ... which the side streams wait for
    # outputs of concurrent replays live in ping-pong buffers: compare the last two frames per stream
    conc = run(models, side)
    for s in range(S):
        for k in (-1, -2):
            assert torch.equal(seq[s][k], conc[s][k]), (s, k)


def test_two_python_threads_on_two_cuda_streams_equal_sequential():
    """re-entrant w.r.t. distinct streams.  The wrapper's bookkeeping that is not per model -- the
    side-stream scope, the registry of deferred tensors, the split-K scratch owner -- is per Python thread, so two
    threads driving two models on two CUDA streams (graph replays; captured beforehand, capture is process-wide in
    torch) produce the same bits as one thread running them one after the other."""
    import threading

    import blockcopy
    from consumers.clips import PolicyFixedFraction, synthetic_clip
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    BS, H, W, T = 64, 256, 512, 6
    clips = [synthetic_clip(T, H, W, seed=20 + s, device="cuda") for s in range(2)]

    def build(s):
        m = blockcopy.BlockCopyModel(build_swiftnet_rn18(seed=0), _settings(block_policy="all", block_size=BS,
                                                                          block_cuda_graphs=True)).eval().cuda().half()
        m.policy = PolicyFixedFraction(BS, fraction=0.3, quantize=2, seed=s)
        return m

    def clip_pass(m, s, keep):
        m.reset_temporal()
        m.policy.reseed(s)
        with torch.no_grad():
            for f in clips[s]:
                o = m(f)
                if keep is not None:
                    keep.append(o.clone())

    models = [build(0), build(1)]
    for s, m in enumerate(models):      # eager pass + capture pass, sequentially
        clip_pass(m, s, None)
        clip_pass(m, s, None)
    seq = [[], []]
    for s, m in enumerate(models):
        clip_pass(m, s, seq[s])
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    par, errors = [[], []], []

    def worker(s):
        try:
            with torch.cuda.stream(streams[s]):
                for _ in range(3):
                    del par[s][:]
                    clip_pass(models[s], s, par[s])
            streams[s].synchronize()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for s in range(2):
        assert len(par[s]) == T
        for a, b in zip(seq[s], par[s]):
            assert torch.equal(a, b)


from side_topologies import SideTopologies as _SideTopologies  # noqa: E402


def test_side_stream_other_topologies_graph_equals_eager():
    import blockcopy
    from consumers.clips import PolicyFixedFraction, deterministic_init_, synthetic_clip

    BS, H, W, T = 32, 128, 256, 6
    clip = synthetic_clip(T, H, W, seed=3, device="cuda")

    def run(graphs):
        torch.manual_seed(0)
        net = deterministic_init_(_SideTopologies().eval(), seed=1)
        m = blockcopy.BlockCopyModel(net, _settings(block_policy="all", block_size=BS, block_cuda_graphs=graphs)).eval().cuda().half()
        m.policy = PolicyFixedFraction(BS, fraction=0.4, quantize=4, seed=5)
        outs = []
        with torch.no_grad():
            for rep in range(3):  # eager, capture, replay
                m.reset_temporal()
                m.policy.reseed(5)
                for t in range(T):
                    o = m(clip[t])
                    if rep == 2:
                        outs.append(o.clone())
        torch.cuda.synchronize()
        return outs

    from blockcopy.core.tensorwrapper import _SideState

    eager = run(False)
    before = _SideState.launches
    graphed = run(True)
    # per eager / captured frame: ds, bn_s+skip, bn_p+pre = 5 kernels on the side stream
    assert _SideState.launches - before >= 5, _SideState.launches - before
    assert all(torch.isfinite(o).all() for o in eager)
    for t, (a, b) in enumerate(zip(eager, graphed)):
        assert torch.equal(a, b), (t, float((a.float() - b.float()).abs().max()))


def test_graph_patch_mode_falls_back_when_a_frame_needs_another_kernel_variant():
    """Whole-frame graphs re-point their input-gather node per frame; a frame whose memory alignment would make
    bc_gather pick another kernel variant than the captured one cannot be patched in: the graph is dropped, the frame
    runs eagerly, results stay those of the eager model, and the next frames capture / replay again."""
    import blockcopy
    from blockcopy.core.argparser import default_settings
    from consumers.clips import PolicyFixedFraction, synthetic_clip
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    H, W, BS = 256, 512, 64
    clip = synthetic_clip(8, H, W, seed=2, dtype=torch.float16, device="cuda")
    odd = torch.empty(3 * H * W + 8, dtype=torch.float16, device="cuda")
    outs = {}
    for graphs in (False, True):
        settings = default_settings(block_policy="all", block_size=BS)
        settings["block_cuda_graphs"] = graphs
        model = blockcopy.BlockCopyModel(build_swiftnet_rn18(seed=1), settings).eval().cuda().half()
        model.policy = PolicyFixedFraction(BS, fraction=0.3, quantize=2, seed=0)
        res = []
        with torch.no_grad():
            for rep in range(3):
                model.reset_temporal()
                for t, f in enumerate(clip):
                    if rep == 2 and t == 5:  # same pixels, storage shifted by one element: 2-byte aligned only
                        f = odd[1:1 + f.numel()].view_as(f).copy_(f)
                    res.append(model(f).clone())
        outs[graphs] = res
        if graphs:
            assert any(k[1] for k in model._graphs.graphs), "graph-patch mode was not used"
    for a, b in zip(outs[False], outs[True]):
        assert torch.equal(a, b)
