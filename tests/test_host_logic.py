"""Host-side logic of the blockcopy package on CPU tensors (kernels rebound to the oracle by
tests/cpu_backend.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden_cpu.py): wrapper state machine, plane bookkeeping, op interception, policies."""
import argparse
import os
import random

import pytest
import torch

from cpu_backend import cpu_backend


def _settings(**kw):
    from blockcopy.core.argparser import default_settings

    return default_settings(**kw)


# ------------------------------------------------------------------------------------------- API surface
def test_public_names():
    import blockcopy

    for n in ("TensorWrapper", "is_block", "is_tensorwrapper", "to_tensorwrapper", "to_tensor", "BlockCopyModel",
              "blockcopy_noblocks", "add_argparser_arguments", "build_policy_from_settings"):
        assert hasattr(blockcopy, n), n
    from blockcopy.utils.profiler import timings  # noqa: F401  consumers import these by path
    from blockcopy.policy.policy import Policy, PolicyAll, PolicyNone, PolicyRandom, PolicyTrainRL  # noqa: F401
    from blockcopy.policy.net import PolicyNet  # noqa: F401
    from blockcopy.utils.block_funcs import CombineFunction, SplitFunction, TransferFunction  # noqa: F401
    from blockcopy.utils.blockpad import BlockPadFunction, pad  # noqa: F401


def test_argparser_flags_and_defaults():
    import blockcopy

    args = vars(blockcopy.add_argparser_arguments(argparse.ArgumentParser()).parse_args([]))
    assert args == dict(block_policy="rl_semseg", block_num_classes=19, block_optim_lr=1e-4, block_optim_wd=1e-3,
                        block_optim_momentum=0, block_target=0.5, block_complexity_weight=5, block_size=128,
                        block_train_interval=4, block_cost_momentum=0.9, block_policy_verbose=False)
    with pytest.raises(NotImplementedError):
        blockcopy.build_policy_from_settings(dict(args, block_policy="static"))
    with pytest.raises(KeyError):  # every key is read even for non-RL policies
        blockcopy.build_policy_from_settings(dict(block_policy="random", block_size=128))


def test_to_tensor_recurses_and_cuda_assert():
    import blockcopy

    t = torch.zeros(1, 1, 2, 2)
    assert blockcopy.to_tensor([t, (t, {"a": t})])[1][1]["a"] is t
    with pytest.raises(AssertionError):
        blockcopy.to_tensorwrapper(t)  # CPU tensor: the block path is CUDA only


def test_timings_api():
    from blockcopy.utils.profiler import Timings

    tm = Timings(level=0)
    with tm.env("a", 1):
        pass
    assert "a" not in tm.records and repr(tm) == "## Profiler: no batches registered"
    tm.set_level(5)
    tm.add_cnt(2)
    with tm.env("a", 1):
        pass
    assert tm.counts["a"] == 1 and "ms per image" in repr(tm)


# ------------------------------------------------------------------------------------------- wrapper semantics
def test_incompatible_and_error_paths():
    import blockcopy
    import torch.nn.functional as F

    with cpu_backend():
        x = blockcopy.to_tensorwrapper(torch.randn(1, 4, 8, 16))
        with pytest.raises(AssertionError):
            x.to_blocks(torch.ones(1, 1, 2, 4, dtype=torch.bool))  # process_temporal_features first
        x.process_temporal_features(None)
        with pytest.raises(AssertionError, match="first run should execute all blocks"):
            x.to_blocks(torch.zeros(1, 1, 2, 4, dtype=torch.bool))
        st = x.process_temporal_features(None)
        b = x.to_blocks(torch.ones(1, 1, 2, 4, dtype=torch.bool))
        assert blockcopy.is_block(b) and b.block_size == 4 and tuple(b.shape) == (8, 4, 4, 4)
        assert not blockcopy.is_block(x) and x.block_size == -1
        for bad in (lambda: F.adaptive_avg_pool2d(b, 1), lambda: b.view(8, -1), lambda: b.flip(0)):
            with pytest.raises(AttributeError, match="not supported for TensorWrapper"):
                bad()
        with pytest.raises(AttributeError, match="already split"):
            b._split(4)
        with pytest.raises(AttributeError, match="Not split in blocks"):
            x.combine()
        with pytest.raises(NotImplementedError, match="equal paddings"):
            F.conv2d(b, torch.randn(4, 4, 3, 3), None, 1, (1, 0))
        y = torch.relu(b) + 1
        assert blockcopy.is_block(y) and y.get_features() is st
        d = y.combine()
        assert blockcopy.is_tensorwrapper(d) and not d.is_blocks and tuple(d.shape) == (1, 4, 8, 16)
        assert type(d.to_tensor()) is torch.Tensor


@pytest.mark.parametrize("channels_last", [False, True])
def test_padded_conv_matches_dense_when_all_blocks_execute(channels_last):
    """With every block executed, halos come from current neighbours, so a padded conv on blocks
    equals the dense conv (no bilinear involved)."""
    import blockcopy
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 4, 16, 24, generator=g)
    w = torch.randn(6, 4, 3, 3, generator=g)
    if channels_last:
        img, w = img.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last)
    with cpu_backend():
        x = blockcopy.to_tensorwrapper(img)
        x.process_temporal_features(None)
        b = x.to_blocks(torch.ones(2, 1, 2, 3, dtype=torch.bool))
        y = F.max_pool2d(F.conv2d(b, w, None, 1, 1), kernel_size=3, stride=2, padding=1)
        out = y.combine().to_tensor()
    ref = F.max_pool2d(F.pad(F.conv2d(img, w, None, 1, 1), (1, 1, 1, 1)), 3, 2, 0)  # zero (not -inf) pool halo
    assert torch.allclose(out, ref, atol=1e-5)


def test_group_norm_folds_blocks_into_one_sample():
    import blockcopy
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(0)
    img = torch.randn(1, 8, 8, 8, generator=g)
    with cpu_backend():
        x = blockcopy.to_tensorwrapper(img)
        x.process_temporal_features(None)
        b = x.to_blocks(torch.ones(1, 1, 2, 2, dtype=torch.bool))
        out = F.group_norm(b, 2).combine().to_tensor()
    assert torch.allclose(out, F.group_norm(img, 2), atol=1e-5)


def test_deconv_weight_packing_is_exact_on_cpu_math():
    """conv2d with the packed weight + depth-to-space == conv_transpose2d (fp64 on the CPU): the phase decomposition
    behind TensorWrapper._try_fused_conv_transpose."""
    import torch.nn.functional as F
    from blockcopy import _C

    g = torch.Generator().manual_seed(0)
    for (k, s, p) in [(4, 2, 1), (4, 4, 0), (2, 2, 0)]:
        x = torch.randn(3, 6, 5, 5, generator=g, dtype=torch.float64)
        w = torch.randn(6, 8, k, k, generator=g, dtype=torch.float64)
        b = torch.randn(8, generator=g, dtype=torch.float64)
        wp, bp = _C.pack_deconv_weight(w, b, s)
        y = F.conv2d(x, wp, bp, padding=wp.shape[2] // 2)
        y = y.reshape(3, s, s, 8, 5, 5).permute(0, 3, 4, 1, 5, 2).reshape(3, 8, 5 * s, 5 * s)
        want = F.conv_transpose2d(x, w, b, stride=s, padding=p)
        assert torch.allclose(y, want, atol=1e-12), (k, s, p)


# ------------------------------------------------------------------------------------------- end to end
def test_swiftnet_clip_matches_reference(golden_dir):
    """SwiftNet-RN18 + BlockCopyModel over a seeded 6-frame clip with replayed masks (incl. an
    empty and a full frame) == the reference's own wrapper on the same weights, fp32 CPU."""
    import blockcopy
    from consumers.clips import PolicyReplay, deterministic_init_, synthetic_clip
    from consumers.swiftnet_rn18 import SwiftNetRN18, fuse_conv_bn_

    fix = torch.load(os.path.join(golden_dir, "swiftnet_cpu_clip.pt"))
    H, W, BS, T = fix["H"], fix["W"], fix["BS"], fix["T"]
    net = deterministic_init_(SwiftNetRN18().eval(), seed=fix["init_seed"])
    model = blockcopy.BlockCopyModel(net, _settings(block_policy="all", block_size=BS)).eval()
    fuse_conv_bn_(model)
    model.policy = PolicyReplay(BS, list(fix["grids"].bool()))
    clip = synthetic_clip(T, H, W, seed=fix["clip_seed"], dtype=torch.float32)
    with cpu_backend(), torch.no_grad():
        model.reset_temporal()
        for t in range(T):
            out = model(clip[t])
            assert type(out) is torch.Tensor and tuple(out.shape) == (1, 19, H // 4, W // 4)
            scale = fix["logits_abs_mean"][t]
            err = (out[:, :, ::4, ::4] - fix["logits_strided"][t]).abs().max().item()
            assert err <= 2e-4 * scale, (t, err, scale)
            agree = (out.argmax(1).to(torch.uint8) == fix["argmax"][t]).float().mean().item()
            assert agree >= 0.9999, (t, agree)
            if t in fix["logits_full"]:
                assert torch.allclose(out, fix["logits_full"][t], atol=2e-4 * scale, rtol=0)
            fs = model.policy_meta["frame_state"]
            assert abs(float(fs.double().sum()) - fix["frame_state_sum"][t]) < 1e-3
            if model.policy_meta["num_exec"] == 0:
                assert out is model.policy_meta["outputs_prev"], "num_exec == 0 returns the previous output object"
    # 21 padded-op planes + 3 combine points (SURVEY.md 3.2 [probe])
    assert len(model.block_temporal_features._planes) == 21
    assert len(model.block_temporal_features._full) == 3


def test_reference_state_dict_names():
    """Consumer model is state_dict-compatible with the reference SwiftNet (names from
    lib/models/swiftnet: backbone.*, spp.spp.*, upsample.N.{bottleneck,blend_conv}.*, logits.*)."""
    from consumers.swiftnet_rn18 import SwiftNetRN18

    keys = set(SwiftNetRN18().state_dict())
    for k in ("backbone.conv1.weight", "backbone.layer2.0.downsample.0.weight", "backbone.layer4.1.bn2.running_var",
              "spp.spp.spp_bn.norm.weight", "spp.spp.spp2.conv.weight", "spp.spp.spp_fuse.conv.weight",
              "upsample.0.bottleneck.conv.weight", "upsample.2.blend_conv.norm.bias", "logits.conv.bias"):
        assert k in keys, k


# ------------------------------------------------------------------------------------------- policies
def test_policy_first_frames_and_quantisation():
    import blockcopy

    x = torch.zeros(1, 3, 512, 1024)
    random.seed(0)
    torch.manual_seed(0)
    pol = blockcopy.build_policy_from_settings(_settings(block_policy="random"))
    meta = pol({"inputs": x, "outputs": None, "outputs_prev": None})
    assert meta["num_exec"] == 32 and meta["grid"].dtype == torch.bool and tuple(meta["grid"].shape) == (1, 1, 4, 8)
    meta.update(outputs=1, outputs_prev=None)
    assert pol(meta)["num_exec"] == 32          # `random` / `none`: the first TWO frames run fully
    meta.update(outputs_prev=1)
    for _ in range(5):
        meta = pol(meta)
        assert meta["num_exec"] % 2 == 0        # rounded up to a multiple of int(32/16)
        assert meta["num_exec"] == int(meta["grid"].sum()) and meta["perc_exec"] == meta["num_exec"] / 32
    none = blockcopy.build_policy_from_settings(_settings(block_policy="none"))
    assert none({"inputs": x, "outputs_prev": 1})["num_exec"] == 0
    with pytest.raises(ZeroDivisionError):      # fewer than 16 blocks: same failure as the reference
        blockcopy.build_policy_from_settings(_settings(block_policy="random"))(
            {"inputs": torch.zeros(1, 3, 256, 256), "outputs_prev": 1})
    with pytest.raises(AssertionError, match="multiple of block size"):
        none({"inputs": torch.zeros(1, 3, 100, 256)})


def test_quantisation_reproduces_reference_choice():
    """random.sample over the skipped cells, seeded: same cells as policy.py:136-143."""
    from blockcopy.policy.policy import PolicyRandom

    pol = PolicyRandom(block_size=128, quantize_number_exec=1 / 16)
    g = torch.Generator().manual_seed(3)
    grid = torch.rand(1, 1, 8, 16, generator=g) < 0.27
    random.seed(11)
    skipped = torch.nonzero(~grid.flatten()).squeeze(1).tolist()
    n = int(grid.sum())
    want = 8 * (1 + (n - 1) // 8)
    expect = grid.clone()
    expect.flatten()[random.sample(skipped, want - n)] = True
    random.seed(11)
    got = pol.quantize_number_exec_grid(grid.clone())
    assert torch.equal(got, expect) and int(got.sum()) == want


def test_policy_net_and_information_gain_match_reference(golden_dir):
    from blockcopy.policy.information_gain import InformationGainSemSeg
    from blockcopy.policy.net import PolicyNet
    from consumers.clips import deterministic_init_

    fix = torch.load(os.path.join(golden_dir, "policy_cpu.pt"))
    g = torch.Generator().manual_seed(fix["input_seed"])
    N, H, W, BS, K = 1, 256, 512, fix["BS"], fix["K"]
    meta = dict(inputs=torch.randn(N, 3, H, W, generator=g), frame_state=torch.randn(N, 3, H, W, generator=g),
                output_repr=torch.randn(N, K, H // 4, W // 4, generator=g),
                grid=torch.rand(N, 1, H // BS, W // BS, generator=g) < 0.4)
    ig_meta = dict(outputs=torch.randn(N, K, H // 4, W // 4, generator=g),
                   outputs_prev=torch.randn(N, K, H // 4, W // 4, generator=g))
    net = deterministic_init_(PolicyNet(block_size=BS, task_num_classes=K), seed=fix["init_seed"]).train()
    assert sum(p.numel() for p in PolicyNet(128, 19).parameters()) == 611211  # SURVEY.md 8(a) a14 [probe]
    with torch.no_grad():
        logits = net(meta)
    assert torch.allclose(logits, fix["logits"], atol=1e-4, rtol=1e-4)
    assert torch.allclose(InformationGainSemSeg(K)(ig_meta), fix["ig"], atol=1e-6, rtol=1e-5)


def test_rl_policy_trains_online():
    """rl_semseg end to end on CPU: Bernoulli masks, REINFORCE step every train_interval frames."""
    import blockcopy
    from consumers.swiftnet_rn18 import build_swiftnet_rn18
    from consumers.clips import synthetic_clip

    random.seed(0)
    torch.manual_seed(0)
    model = blockcopy.BlockCopyModel(build_swiftnet_rn18(), _settings(block_policy="rl_semseg", block_size=64,
                                                                    block_train_interval=2, block_optim_lr=1e-3)).eval()
    model.policy.net.train()
    before = [p.detach().clone() for p in model.policy.net.parameters()]
    clip = synthetic_clip(5, 256, 512, seed=1, dtype=torch.float32)
    with cpu_backend(), torch.no_grad():
        model.reset_temporal()
        for f in clip:
            out = model(f)
    meta = model.policy_meta
    for k in ("inputs", "outputs", "outputs_prev", "grid", "num_exec", "num_total", "perc_exec", "frame_state",
              "output_repr", "grid_log_probs", "grid_probs", "information_gain"):
        assert k in meta, k
    assert tuple(meta["information_gain"].shape) == (1, 1, 16, 32) and tuple(out.shape) == (1, 19, 64, 128)
    assert any(not torch.equal(a, b) for a, b in zip(before, model.policy.net.parameters())), "policy did not train"
    assert 0 < model.policy.stats.get_exec_percentage() <= 1


def test_lazy_fusion_removes_scatters_and_elementwise_kernels(monkeypatch):
    """Bookkeeping of the deferred-epilogue engine (kernels emulated on CPU): with fusion on, convs
    write the next op's plane themselves and ReLU / add / BN / bilinear never run as separate ops."""
    import blockcopy
    from blockcopy import _C
    from blockcopy.core import tensorwrapper as tw
    from consumers.clips import PolicyReplay, deterministic_init_, synthetic_clip
    from consumers.swiftnet_rn18 import SwiftNetRN18, fuse_conv_bn_

    grids = [torch.ones(1, 1, 4, 8, dtype=torch.bool), torch.rand(1, 1, 4, 8, generator=torch.Generator().manual_seed(0)) < 0.4]
    clip = synthetic_clip(2, 256, 512, seed=1, dtype=torch.float32)

    def run(lazy):
        monkeypatch.setattr(tw, "LAZY_FUSION", lazy)
        net = deterministic_init_(SwiftNetRN18().eval(), seed=0)
        model = blockcopy.BlockCopyModel(net, _settings(block_policy="all", block_size=64)).eval()
        fuse_conv_bn_(model)
        model.policy = PolicyReplay(64, grids)
        counts = {}
        with cpu_backend(), torch.no_grad():
            for name in ("scatter", "conv_igemm", "ew_fused", "gather_halo", "head_1x1"):
                orig = getattr(_C, name)
                def counted(*a, _o=orig, _n=name, **k):
                    counts[_n] = counts.get(_n, 0) + 1
                    return _o(*a, **k)
                setattr(_C, name, counted)
            model.reset_temporal()
            outs = [model(f).clone() for f in clip]
        return outs, counts

    eager, c0 = run(False)
    lazy, c1 = run(True)
    for a, b in zip(eager, lazy):
        assert torch.allclose(a, b, atol=1e-3 * float(a.abs().mean()))
    assert c0["conv_igemm"] == c1["conv_igemm"] == 2 * 25       # 20 3x3 + 3 downsample + 3 skip 1x1 - (logits: Cout 19)
    # 3 skip BN-ReLU, 3 upsample+add+BN+ReLU, 4 dense BN-ReLU in the SPP (254 ch: torch); the logits' BN-ReLU is
    # absorbed, with the 19-channel conv and the final combine, into one bc_head_1x1 launch per frame
    assert c1["ew_fused"] == 2 * 10 and c1["head_1x1"] == 2 and "head_1x1" not in c0, (c0, c1)
    assert c0["scatter"] - c1["scatter"] >= 2 * 18, (c0, c1)     # planes are written by producer epilogues instead


def test_policy_will_train_hint_is_set_and_a_wrong_hint_fails_loudly():
    """BlockCopyModel announces training frames to the policy (policy_meta['policy_will_train']); a training step
    on a frame that was announced as inference-only raises instead of silently skipping the update."""
    import pytest
    import torch

    import blockcopy
    from blockcopy.core.argparser import default_settings
    from blockcopy.policy.policy import Policy

    seen = []

    class Spy(Policy):
        def forward(self, policy_meta):
            seen.append(policy_meta.get("policy_will_train"))
            shape = self._grid_shape(policy_meta)
            policy_meta["grid"] = torch.ones(shape, dtype=torch.bool)
            return self.stats.add_policy_meta(policy_meta)

    with cpu_backend():
        model = blockcopy.BlockCopyModel(torch.nn.Conv2d(3, 4, 1), default_settings(block_policy="all", block_size=8,
                                                                                  block_train_interval=3)).eval()
        model.policy = Spy(block_size=8)
        x = torch.randn(1, 3, 16, 16)
        with torch.no_grad():
            for _ in range(6):
                model(x)
    assert seen == [False, False, True, False, False, True]

    pol = blockcopy.build_policy_from_settings(default_settings(block_policy="rl_semseg", block_size=8))
    meta = dict(inputs=x, outputs=torch.randn(1, 19, 4, 4), outputs_prev=torch.randn(1, 19, 4, 4),
                grid=torch.ones(1, 1, 2, 2, dtype=torch.bool), perc_exec=1.0, grid_log_probs=None,
                grid_probs=torch.full((1, 1, 2, 2), 0.5))
    with pytest.raises(RuntimeError, match="policy_will_train"):
        pol.optim(meta, train=True)


def test_deferred_engine_on_other_topologies_equals_op_by_op(monkeypatch):
    """Lazy epilogue fusion + tile-less launches (kernels emulated on CPU, unwritten tile batches poisoned with NaN)
    against op-by-op execution on a model with main-chain 1x1 convs, a pre-activation unit feeding a padded conv,
    results with two consumers, an in-place op, and an upsampled sum read by two padded convs and a plain add."""
    import blockcopy
    from blockcopy.core import tensorwrapper as tw
    from consumers.clips import PolicyFixedFraction, deterministic_init_, synthetic_clip
    from side_topologies import SideTopologies

    clip = synthetic_clip(4, 128, 256, seed=3, dtype=torch.float32)

    def run(lazy):
        monkeypatch.setattr(tw, "LAZY_FUSION", lazy)
        net = deterministic_init_(SideTopologies().eval(), seed=1)
        model = blockcopy.BlockCopyModel(net, _settings(block_policy="all", block_size=32)).eval()
        model.policy = PolicyFixedFraction(32, fraction=0.4, quantize=4, seed=5)
        with cpu_backend(), torch.no_grad():
            model.reset_temporal()
            return [model(f).clone() for f in clip]

    lazy, plain = run(True), run(False)
    for t, (a, b) in enumerate(zip(lazy, plain)):
        assert torch.isfinite(a).all(), f"frame {t}: an unwritten tile batch was read"
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4), (t, float((a - b).abs().max()))


def test_num_exec_hint_is_ignored_after_the_grid_was_edited():
    """the executed-block count attached to a grid is only honoured while the tensor is unchanged
    (in-place version counter); an edited grid is counted again like the reference does (tensorwrapper.py:157)."""
    import blockcopy
    from blockcopy.utils.hints import get_num_exec_hint, set_num_exec_hint

    grid = torch.ones(1, 1, 2, 4, dtype=torch.bool)
    set_num_exec_hint(grid, 8)
    assert get_num_exec_hint(grid) == 8
    with cpu_backend():
        x = blockcopy.to_tensorwrapper(torch.randn(1, 4, 8, 16))
        st = x.process_temporal_features(None)
        b = x.to_blocks(grid)
        assert tuple(b.shape) == (8, 4, 4, 4)
        b.combine_()
        g2 = torch.ones(1, 1, 2, 4, dtype=torch.bool)
        set_num_exec_hint(g2, 8)
        g2[0, 0, 0, 1] = False          # a custom policy edits its grid after the count was taken
        g2[0, 0, 1, 2] = False
        assert get_num_exec_hint(g2) is None
        x2 = blockcopy.to_tensorwrapper(torch.randn(1, 4, 8, 16))
        x2.process_temporal_features(st)
        b2 = x2.to_blocks(g2)
        assert tuple(b2.shape) == (6, 4, 4, 4) and b2.get_mapping_exec().tolist() == [0, 2, 3, 4, 5, 7]


def test_object_detection_information_gain_host_logic_vs_reference_fixture(golden_dir):
    """Host side of InformationGainObjectDetection (IoU matching, value arithmetic, Python-slice box semantics)
    against the masks of the unmodified reference (tests/golden/det_ig_kat.npz, made on a B200 because the
    reference hard-wires device 'cuda' there); the device rasteriser is replaced by numpy slice painting."""
    import numpy as np

    from blockcopy.policy import information_gain as IG

    path = os.path.join(golden_dir, "det_ig_kat.npz")
    if not os.path.exists(path):
        pytest.skip("det_ig_kat.npz not generated yet")
    fix = np.load(path)
    H, W, T = int(fix["H"]), int(fix["W"]), int(fix["n_frames"])

    def paint(out2d, rects, values, shift):
        o = np.zeros(tuple(out2d.shape), np.float32)
        for (x1, y1, x2, y2), v in zip(rects, values):
            ys, xs = slice(y1 << shift, y2 << shift), slice(x1 << shift, x2 << shift)
            o[ys, xs] = np.maximum(o[ys, xs], v)
        out2d.copy_(torch.from_numpy(o))

    saved = IG.InformationGainObjectDetection._paint
    IG.InformationGainObjectDetection._paint = staticmethod(paint)
    try:
        ig = IG.InformationGainObjectDetection(num_classes=1)
        inputs = torch.zeros(1, 3, H, W)
        for t in range(T):
            meta = dict(inputs=inputs, outputs=[[fix[f"boxes_{t}"]]],
                        outputs_prev=[[fix[f"boxes_{t - 1}"]]] if t else None)
            assert np.array_equal(ig.get_output_repr(meta).numpy(), fix[f"repr_{t}"]), t
            if t:
                assert np.array_equal(ig(meta).numpy(), fix[f"gain_{t}"]), t
    finally:
        IG.InformationGainObjectDetection._paint = saved


def test_non_inplace_combine_reuses_the_previous_plane_only_when_nobody_can_see_it():
    """combine() (non in place, reference tensorwrapper.py:421-434: clone + scatter) may update the stored previous
    tensor when the store is its only owner -- the clone is unobservable then; as soon as anybody holds the old
    tensor, a view or an alias of it, the copy is made and the old tensor keeps its values."""
    import blockcopy

    g = torch.Generator().manual_seed(0)
    full = torch.ones(1, 1, 2, 2, dtype=torch.bool)
    part = torch.tensor([[[[True, False], [False, True]]]])
    frames = [torch.randn(1, 8, 8, 8, generator=g) for _ in range(4)]
    with cpu_backend():
        def step(frame, grid, state):
            x = blockcopy.to_tensorwrapper(frame.clone())
            state = x.process_temporal_features(state)
            return x.to_blocks(grid).combine().to_tensor(), state

        out0, st = step(frames[0], full, None)
        ptr0, keep0 = out0.data_ptr(), out0.clone()
        # (a) the caller still holds frame 0's result: frame 1 must not touch it
        out1, st = step(frames[1], part, st)
        assert out1.data_ptr() != ptr0 and torch.equal(out0, keep0)
        want1 = frames[0].clone()
        want1[..., :4, :4], want1[..., 4:, 4:] = frames[1][..., :4, :4], frames[1][..., 4:, 4:]
        assert torch.equal(out1, want1)
        # (b) a view of the old result is still alive
        ptr1, view1 = out1.data_ptr(), out1[:, :2]
        del out1
        out2, st = step(frames[2], part, st)
        assert out2.data_ptr() != ptr1 and torch.equal(view1, want1[:, :2])
        # (c) nothing refers to frame 2's result any more: frame 3 is combined into the same storage
        ptr2 = out2.data_ptr()
        want3 = out2.clone()
        want3[..., :4, :4], want3[..., 4:, 4:] = frames[3][..., :4, :4], frames[3][..., 4:, 4:]
        del out2, view1
        out3, st = step(frames[3], part, st)
        assert out3.data_ptr() == ptr2 and torch.equal(out3, want3)


def test_policy_checks_on_cpu_are_immediate_and_deferred_queue_is_bounded(capsys):
    """PolicyTrainRL._check: on CPU tensors (no sync to save) the reference's in-place behaviour -- AssertionError, or a
    printed warning -- and flush_checks() on an empty queue is a no-op."""
    import blockcopy

    policy = blockcopy.build_policy_from_settings(_settings(block_policy="rl_semseg", block_size=64))
    policy._check(torch.tensor(False), "never shown")
    with pytest.raises(AssertionError, match="boom"):
        policy._check(torch.tensor(True), "boom")
    policy._check(torch.tensor(True), "only a warning", warn=True)
    assert "only a warning" in capsys.readouterr().out
    assert policy._deferred == []
    policy.flush_checks()
    assert policy._read_count_and_checks(torch.tensor([7, 5], dtype=torch.int32)) == 7
