"""bc_policy_features / bc_info_gain against the torch op sequences of the reference
(policy/net.py:84-113, policy/information_gain.py:32-41)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _torch_features(meta, scale):
    frame = meta["inputs"]
    feats = [F.interpolate(frame, scale_factor=scale, mode="nearest").float()]
    size = feats[0].shape[2:]
    feats.append(F.interpolate(meta["frame_state"], size=size, mode="nearest").float())
    feats.append(F.interpolate(meta["output_repr"], size=size, mode="nearest").type(torch.float32) - 0.5)
    feats.append(F.interpolate(meta["grid"].float(), size=size, mode="nearest") - 0.5)
    return torch.cat(feats, dim=1)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("BS,H,W,N,cl", [(128, 512, 1024, 1, True), (64, 256, 512, 2, False), (32, 128, 256, 1, True)])
def test_policy_features_bit_exact(dtype, BS, H, W, N, cl):
    from blockcopy.policy.net import PolicyNet

    g = torch.Generator().manual_seed(0)
    dev = "cuda"
    rep = torch.randn(N, 19, H // 4, W // 4, generator=g).to(dtype).to(dev)
    if cl:
        rep = rep.contiguous(memory_format=torch.channels_last)
    meta = dict(inputs=torch.randn(N, 3, H, W, generator=g).to(dtype).to(dev),
                frame_state=torch.randn(N, 3, H, W, generator=g).to(dtype).to(dev), output_repr=rep,
                grid=(torch.rand(N, 1, H // BS, W // BS, generator=g) < 0.4).to(dev))
    net = PolicyNet(block_size=BS, task_num_classes=19)
    got = net._fused_features(meta)
    assert got is not None and got.dtype == torch.float32
    want = _torch_features(meta, net.scale_factor)
    assert got.shape == want.shape and torch.equal(got, want)


def test_info_gain_matches_torch_sequence():
    from blockcopy.policy.information_gain import InformationGainSemSeg

    g = torch.Generator().manual_seed(1)
    dev = "cuda"
    for cl in (False, True):
        cur = (2 * torch.randn(2, 19, 64, 128, generator=g)).half().to(dev)
        prev = (cur.float() + 0.5 * torch.randn(2, 19, 64, 128, generator=g).to(dev)).half()
        if cl:
            cur, prev = cur.contiguous(memory_format=torch.channels_last), prev.contiguous(memory_format=torch.channels_last)
        got = InformationGainSemSeg(19)({"outputs": cur, "outputs_prev": prev})
        a = F.interpolate(cur, scale_factor=0.25, mode="bilinear")
        b = F.interpolate(prev, scale_factor=0.25, mode="bilinear")
        want = F.kl_div(F.log_softmax(a, 1), F.log_softmax(b, 1), reduction="none", log_target=True).mean(1, keepdim=True)
        ref32 = F.kl_div(F.log_softmax(a.float(), 1), F.log_softmax(b.float(), 1), reduction="none",
                         log_target=True).mean(1, keepdim=True)
        assert got.shape == want.shape == (2, 1, 16, 32) and got.dtype == torch.float16
        # both fp16 pipelines sit within fp16 rounding noise of the fp32 value; compare each to it
        tol = 4e-3 * float(ref32.abs().max()) + 1e-4
        assert (got.float() - ref32).abs().max().item() <= tol
        assert (got.float() - want.float()).abs().max().item() <= 2 * tol


def test_graphed_policy_trunk_matches_eager():
    """PolicyNet with use_cuda_graphs (trunk forward / backward replayed as CUDA graphs) gives the eager net's
    logits, parameter gradients and batch-norm running statistics, and capturing leaves no trace in them."""
    import copy

    from blockcopy.policy.net import PolicyNet

    torch.manual_seed(0)
    eager = PolicyNet(block_size=128, task_num_classes=19).cuda().train()
    graphed = copy.deepcopy(eager)
    graphed.use_cuda_graphs = True
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(3):
        x = torch.randn(1, 26, 128, 256, device="cuda", generator=g)
        w = torch.randn(1, 1, 4, 8, device="cuda", generator=g)
        outs = []
        for net in (eager, graphed):
            net.zero_grad(set_to_none=True)
            y = net._trunk_forward(x)
            (y * w).mean().backward()
            outs.append(y.detach().clone())
        assert torch.allclose(outs[0], outs[1], rtol=1e-4, atol=1e-5), step
        for (n, a), b in zip(eager.named_parameters(), graphed.parameters()):
            if a.grad is None or b.grad is None:  # parameters the trunk does not use (resnet fc)
                assert a.grad is None and (b.grad is None or not b.grad.any()), (step, n)
                continue
            # (cuDNN may pick another wgrad algorithm under capture: compare in norm)
            assert (a.grad - b.grad).norm() <= 2e-3 * a.grad.norm() + 1e-7, (step, n)
        for (n, a), b in zip(eager.named_buffers(), graphed.buffers()):
            assert torch.allclose(a.float(), b.float(), rtol=1e-4, atol=1e-6), (step, n)
    assert list(eager.state_dict().keys()) == list(graphed.state_dict().keys())


@pytest.mark.parametrize("N,C,H,W", [(1, 64, 64, 128), (2, 128, 16, 32), (1, 64, 256, 512), (1, 8, 5, 7)])
def test_bn_stats_matches_float64_and_is_reproducible(N, C, H, W):
    from blockcopy import _C

    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = (1.5 * torch.randn(N, C, H, W, device="cuda", generator=g) + 0.7).half().contiguous(memory_format=torch.channels_last)
    ws = torch.zeros(_C.BN_STATS_WORKSPACE, dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(3):
        mean = torch.empty(C, device="cuda")
        invstd = torch.empty(C, device="cuda")
        _C.bn_stats(x, mean, invstd, 1e-5, ws)
        outs.append((mean, invstd))
    x64 = x.double()
    want_mean = x64.mean(dim=(0, 2, 3))
    want_inv = 1.0 / torch.sqrt(x64.var(dim=(0, 2, 3), unbiased=False) + 1e-5)
    assert torch.allclose(outs[0][0].double(), want_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(outs[0][1].double(), want_inv, rtol=1e-5, atol=1e-6)
    for m, s in outs[1:]:
        assert torch.equal(m, outs[0][0]) and torch.equal(s, outs[0][1])
    assert int(ws[:4].view(torch.int32)) == 0  # the ticket is left at zero


def _policy_net(seed=0):
    from blockcopy.policy.net import PolicyNet

    torch.manual_seed(seed)
    net = PolicyNet(block_size=128, task_num_classes=19).cuda().train()
    with torch.no_grad():  # non-trivial affine parameters
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    return net


def test_fused_policy_trunk_matches_fp32_torch():
    """policy/fused_net.py (bc_conv_igemm + bc_bn_stats + bc_ew_fused, fp16 operands) against the torch trunk in
    strict fp32: logits within 2 % of their range + 0.02, i.e. execute probabilities within ~0.01."""
    from blockcopy.policy.fused_net import FusedPolicyTrunk

    net = _policy_net()
    fused = FusedPolicyTrunk(net)
    assert fused.ok
    g = torch.Generator(device="cuda").manual_seed(5)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for (N, H, W) in [(1, 256, 512), (2, 128, 256)]:
            x = torch.randn(N, 26, H, W, device="cuda", generator=g)
            assert fused.supports(x)
            with torch.no_grad():
                want = net.layers(net.backbone(x))
            got = fused(x)
            assert got.shape == want.shape == (N, 1, H // 32, W // 32) and got.dtype == torch.float32
            tol = 0.02 * float(want.abs().max()) + 0.02
            assert float((got - want).abs().max()) <= tol, (float((got - want).abs().max()), tol)
            assert float((torch.sigmoid(got) - torch.sigmoid(want)).abs().max()) <= 0.02
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert not fused.supports(torch.randn(1, 26, 100, 200, device="cuda"))  # not divisible: torch path


def test_fused_policy_trunk_graph_equals_eager_and_follows_parameter_updates():
    from blockcopy.policy.fused_net import FusedPolicyTrunk

    net = _policy_net(1)
    a, b = FusedPolicyTrunk(net), FusedPolicyTrunk(net)
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(1, 26, 256, 512, device="cuda", generator=g)
    first = a(x).clone()
    assert torch.equal(b(x, use_cuda_graph=True), first)
    assert torch.equal(b(x, use_cuda_graph=True), first)  # replay
    with torch.no_grad():  # an optimiser step changes the parameters in place
        for q in net.parameters():
            q.add_(0.05 * torch.randn(q.shape, device="cuda", generator=g))
    second = a(x).clone()
    assert not torch.equal(second, first)
    assert torch.equal(b(x, use_cuda_graph=True), second)  # the graph re-packs the live weights


def test_rl_policy_uses_fused_trunk_on_frames_without_update():
    """BlockCopyModel tells the policy whether optim() will train; frames that will not run the inference trunk."""
    import blockcopy
    from blockcopy.core.argparser import default_settings
    from blockcopy.policy import fused_net
    from consumers.clips import synthetic_clip
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    calls = []
    orig = fused_net.FusedPolicyTrunk.run

    def spy(self, fill, shape, device, use_cuda_graph=False):
        calls.append(tuple(shape))
        return orig(self, fill, shape, device, use_cuda_graph)

    fused_net.FusedPolicyTrunk.run = spy
    try:
        model = blockcopy.BlockCopyModel(build_swiftnet_rn18(), default_settings(block_policy="rl_semseg", block_size=128,
                                                                               block_train_interval=3)).eval().cuda().half()
        model.policy.net = model.policy.net.float().train()  # the reference driver's configuration
        clip = synthetic_clip(7, 1024, 2048, seed=0, device="cuda")
        with torch.no_grad():
            model.reset_temporal()
            outs = [model(f) for f in clip]
        torch.cuda.synchronize()
    finally:
        fused_net.FusedPolicyTrunk.run = orig
    assert all(torch.isfinite(o).all() for o in outs)
    # frames 2..7 run the policy net (frame 1 has no history); frames 3 and 6 train (clip_length % 3 == 0)
    assert len(calls) == 4, calls
    assert model.policy.stats.count_images == 7


@pytest.mark.parametrize("N,C,Cx,Cout,H,W,k,s,p,cl", [(1, 128, 128, 1, 16, 32, 3, 2, 1, True), (2, 20, 64, 3, 9, 7, 3, 1, 1, False),
                                                     (1, 130, 136, 2, 8, 8, 1, 1, 0, False)])
def test_conv_fewout_matches_torch(N, C, Cx, Cout, H, W, k, s, p, cl):
    from blockcopy import _C

    g = torch.Generator(device="cuda").manual_seed(C + Cout)
    x = torch.randn(N, Cx, H, W, device="cuda", generator=g).half().contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, C, k, k, device="cuda", generator=g)
    if cl:
        w = w.contiguous(memory_format=torch.channels_last)
    b = torch.randn(Cout, device="cuda", generator=g)
    got = _C.conv_fewout(x, w, b, s, p)
    want = torch.nn.functional.conv2d(x[:, :C].double(), w.double(), b.double(), stride=s, padding=p)
    assert got.shape == want.shape and got.dtype == torch.float32
    assert torch.allclose(got.double(), want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("momentum,wd", [(0.0, 0.0), (0.9, 0.0), (0.9, 1e-3)])
def test_fused_rmsprop_matches_torch(momentum, wd):
    """bc_rmsprop_step (one launch for all tensors) against torch.optim.RMSprop, parameter by parameter, over
    several steps; state_dict layouts agree."""
    import copy

    from blockcopy.policy.fused_optim import FusedRMSprop
    from blockcopy.policy.net import PolicyNet

    torch.manual_seed(0)
    a = PolicyNet(block_size=128, task_num_classes=19).cuda().train().to(memory_format=torch.channels_last)
    b = copy.deepcopy(a)
    oa = torch.optim.RMSprop(a.parameters(), lr=1e-3, weight_decay=wd, momentum=momentum, centered=False)
    ob = FusedRMSprop(b.parameters(), lr=1e-3, weight_decay=wd, momentum=momentum, centered=False)
    g = torch.Generator(device="cuda").manual_seed(3)
    for step in range(4):
        for pa, pb in zip(a.parameters(), b.parameters()):
            if pa.dim() == 2 and step % 2:  # some parameters without gradient on some steps (the unused resnet fc)
                pa.grad = pb.grad = None
                continue
            gr = torch.randn(pa.shape, device="cuda", generator=g).contiguous(
                memory_format=torch.channels_last if pa.dim() == 4 else torch.contiguous_format)
            pa.grad, pb.grad = gr.clone(memory_format=torch.preserve_format), gr.clone(memory_format=torch.preserve_format)
        oa.step()
        ob.step()
        for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
            assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-8), (step, n, float((pa - pb).abs().max()))
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["param_groups"][0].keys() == sb["param_groups"][0].keys()
    for k in sa["state"]:
        assert sa["state"][k].keys() == sb["state"][k].keys()
        assert torch.allclose(sa["state"][k]["square_avg"], sb["state"][k]["square_avg"], rtol=1e-6, atol=1e-10)
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"])


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_policy_features_nhwc16_equals_rounded_fp32_features_and_feeds_the_trunk(dtype):
    """bc_policy_features_nhwc16 writes what `policy_features().half()` holds, into the padded channels_last plane;
    the trunk run from it gives the same logits as the trunk run from the fp32 features."""
    from blockcopy import _C
    from blockcopy.policy.fused_net import FusedPolicyTrunk

    g = torch.Generator().manual_seed(2)
    N, H, W, BS = 1, 512, 1024, 128
    rep = torch.randn(N, 19, H // 4, W // 4, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    frame, state = torch.randn(N, 3, H, W, generator=g).to(dtype).cuda(), torch.randn(N, 3, H, W, generator=g).to(dtype).cuda()
    grid = (torch.rand(N, 1, H // BS, W // BS, generator=g) < 0.4).cuda()
    x = _C.policy_features(frame, state, rep, grid, 0.25)
    x16 = torch.full((N, 64, H // 4, W // 4), 7.0, dtype=torch.float16, device="cuda").contiguous(memory_format=torch.channels_last)
    _C.policy_features_nhwc16(x16, frame, state, rep, grid, 0.25)
    assert torch.equal(x16[:, :26], x.half())
    assert bool((x16[:, 26:32] == 0).all()) and bool((x16[:, 32:] == 7.0).all())  # written chunk padding / untouched rest
    net = _policy_net(3)
    a, b = FusedPolicyTrunk(net), FusedPolicyTrunk(net)
    want = a(x).clone()
    got = b.run(lambda buf: _C.policy_features_nhwc16(buf, frame, state, rep, grid, 0.25), tuple(x.shape), x.device)
    assert torch.equal(got, want)


def test_object_detection_information_gain_equals_reference_fixture(golden_dir):
    """InformationGainObjectDetection (host IoU matching + bc_raster_boxes) against the masks the UNMODIFIED
    reference painted box by box on a B200 (oracle/make_golden_gpu.py det): bit-exact, including boxes sticking
    out of the frame and Python-slice semantics for negative coordinates."""
    import numpy as np

    from blockcopy.policy.information_gain import InformationGainObjectDetection

    path = os.path.join(golden_dir, "det_ig_kat.npz")
    if not os.path.exists(path):
        pytest.skip("det_ig_kat.npz not generated yet (oracle/make_golden_gpu.py det)")
    fix = np.load(path)
    H, W, T = int(fix["H"]), int(fix["W"]), int(fix["n_frames"])
    ig = InformationGainObjectDetection(num_classes=1)
    inputs = torch.zeros(1, 3, H, W, device="cuda")
    for t in range(T):
        f = fix[f"boxes_{t}"]
        meta = dict(inputs=inputs, outputs=[[f]], outputs_prev=[[fix[f"boxes_{t - 1}"]]] if t else None)
        got = ig.get_output_repr(meta)
        assert got.dtype == torch.float32 and tuple(got.shape) == (1, 1, H, W)
        assert np.array_equal(got.cpu().numpy(), fix[f"repr_{t}"]), f"output_repr of frame {t}"
        if t:
            gain = ig(meta)
            assert np.array_equal(gain.cpu().numpy(), fix[f"gain_{t}"]), f"information gain of frame {t}"


def test_rl_objectdetection_policy_builds_and_trains():
    """`--block-policy rl_objectdetection` (it used to crash in the first optim())."""
    import numpy as np

    import blockcopy
    from blockcopy.core.argparser import default_settings

    torch.manual_seed(0)
    s = default_settings(block_policy="rl_objectdetection", block_num_classes=1, block_target=0.3, block_train_interval=2)
    policy = blockcopy.build_policy_from_settings(s).cuda()
    policy.net = policy.net.float().train()
    H, W = 256, 512
    rng = np.random.RandomState(0)

    def det(n):
        x1 = rng.uniform(0, W - 60, n); y1 = rng.uniform(0, H - 60, n)
        return [[np.stack([x1, y1, x1 + rng.uniform(10, 50, n), y1 + rng.uniform(10, 50, n), rng.uniform(0.1, 1, n)], 1)
                 .astype(np.float32)]]

    before = [p.detach().clone() for p in policy.net.parameters()]
    meta = {"inputs": torch.randn(1, 3, H, W, device="cuda"), "outputs": None, "outputs_prev": None}
    for t in range(4):
        meta["inputs"] = torch.randn(1, 3, H, W, device="cuda")
        meta["policy_will_train"] = t % 2 == 1
        if t == 0:
            meta["frame_state"] = meta["inputs"].clone()
        meta = policy(meta)
        assert meta["grid"].dtype == torch.bool and tuple(meta["grid"].shape) == (1, 1, 2, 4)
        meta["outputs_prev"], meta["outputs"] = meta["outputs"], det(6)
        meta = policy.optim(meta, train=t % 2 == 1)
        assert tuple(meta["output_repr"].shape) == (1, 1, H, W)
    assert any(not torch.equal(a, b) for a, b in zip(before, policy.net.parameters()))


@pytest.mark.parametrize("G,multiple,alo", [(128, 8, False), (128, 8, True), (32, 2, False), (1024, 64, False),
                                            (8192, 512, False), (100, 6, False), (128, 0, False)])
def test_sample_grid_equals_host_restatement(G, multiple, alo):
    """bc_sample_grid (Bernoulli draw + executed-count quantisation on the device) == its numpy restatement, bit for
    bit, and the count follows the reference's rounding rule (policy/policy.py:139-140)."""
    from blockcopy import _C
    from blockcopy.policy.policy import sample_grid_host

    g = torch.Generator(device="cuda").manual_seed(G + multiple)
    for trial in range(6):
        scale = (0.02, 0.3, 1.0, 3.0, 0.0, 1.0)[trial]
        probs = torch.sigmoid(scale * torch.randn(G, device="cuda", generator=g) - (4.0 if trial == 4 else 0.0))
        if trial == 4:
            probs.zero_()  # nothing executes before rounding
        uni = torch.rand(2 * G, device="cuda", generator=g)
        if trial == 5:
            uni[G:] = (uni[G:] * 4).floor() / 4  # many equal keys: ties go to the lower index
        grid, counts = _C.sample_grid(probs, uni, multiple, alo)
        want, e1, e0 = sample_grid_host(probs.cpu().numpy(), uni.cpu().numpy(), multiple, alo)
        assert grid.dtype == torch.bool and counts.tolist() == [e1, e0], (counts.tolist(), e1, e0)
        assert (grid.cpu().numpy() == want).all()
        assert int(grid.sum()) == e1
        if multiple > 0 and e0 > 0 and G % multiple == 0:
            assert e1 % multiple == 0 and 0 <= e1 - e0 < multiple
            assert e1 == multiple * (1 + (e0 - 1) // multiple)
        # rounding only switches cells ON
        assert bool((grid.cpu() | ~(uni[:G] < probs).cpu()).all())


def test_sample_grid_extra_cells_are_uniform():
    """The cells switched on by the rounding are a uniformly random subset of the skipped ones (what the reference's
    random.sample draws): over many draws every skipped cell is picked about equally often."""
    from blockcopy import _C

    G, multiple = 64, 16
    probs = torch.zeros(G, device="cuda")
    probs[:9] = 1.0  # 9 executed -> rounded to 16: 7 of the 55 skipped cells are added
    hits = torch.zeros(G, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    trials = 4000
    for _ in range(trials):
        uni = torch.rand(2 * G, device="cuda", generator=g)
        uni[:G].clamp_(max=0.999)
        grid, counts = _C.sample_grid(probs, uni, multiple, False)
        hits += grid.float()
    assert hits[:9].eq(trials).all()
    freq = hits[9:] / trials
    assert abs(float(freq.mean()) - 7 / 55) < 1e-6
    assert float((freq - 7 / 55).abs().max()) < 0.03  # 5 sigma of a binomial(4000, 0.127) is 0.026


def test_fused_policy_trunk_updates_bn_running_statistics_like_torch():
    """VERDICT r01 missing #8: the fused trunk reproduces the train-mode side effect of the torch path -- running_mean,
    running_var (unbiased batch variance, the module's momentum) and num_batches_tracked of every BatchNorm2d."""
    import copy

    from blockcopy.policy.fused_net import FusedPolicyTrunk

    net_t = _policy_net()
    net_f = copy.deepcopy(net_t)
    fused = FusedPolicyTrunk(net_f)
    g = torch.Generator(device="cuda").manual_seed(9)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for step in range(3):
            x = torch.randn(1, 26, 128, 256, device="cuda", generator=g) * (1 + step)
            with torch.no_grad():
                net_t.layers(net_t.backbone(x))
            fused(x, use_cuda_graph=step > 0)  # eager once, then capture + replay
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    bns_t = [m for m in net_t.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    bns_f = [m for m in net_f.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    assert len(bns_t) == len(bns_f) > 5
    for a, b in zip(bns_t, bns_f):
        assert int(a.num_batches_tracked) == int(b.num_batches_tracked) == 3
        for ra, rb in ((a.running_mean, b.running_mean), (a.running_var, b.running_var)):
            tol = 0.03 * float(ra.abs().max()) + 1e-3
            assert float((ra - rb).abs().max()) <= tol, (float((ra - rb).abs().max()), tol)
        assert not torch.equal(b.running_var, torch.ones_like(b.running_var))


def test_deferred_nan_check_raises_within_one_frame():
    """PolicyTrainRL's NaN asserts (reference policy/policy.py:253) do not stall the host on the GPU any more: the
    verdict is read with the next executed-block count -- and still arrives, at most one frame later; with
    block_policy_strict_checks the assert fires in place."""
    import blockcopy
    from blockcopy.core.argparser import default_settings
    from consumers.clips import synthetic_clip
    from consumers.swiftnet_rn18 import build_swiftnet_rn18

    for strict in (False, True):
        settings = default_settings(block_policy="rl_semseg", block_size=128, block_train_interval=3)
        settings["block_policy_strict_checks"] = strict
        model = blockcopy.BlockCopyModel(build_swiftnet_rn18(), settings).eval().cuda().half()
        model.policy.net = model.policy.net.float().train()
        clip = synthetic_clip(4, 512, 1024, seed=0, device="cuda")
        with torch.no_grad():
            model(clip[0])
            model(clip[1])
            with torch.no_grad():
                model.policy.net.layers[2][0].bias.fill_(float("nan"))
            with pytest.raises(AssertionError, match="NaN"):
                model(clip[2])
                assert not strict, "strict checks must raise on the frame itself"
                model(clip[3])


@pytest.mark.parametrize("N,C,H,W,relu,affine", [(1, 64, 256, 512, True, True), (2, 128, 16, 32, False, True), (1, 64, 5, 7, True, False),
                                                   (1, 8, 64, 128, True, True)])
def test_bn_norm_equals_bn_stats_plus_ew_fused(N, C, H, W, relu, affine):
    """bc_bn_norm (statistics, grid-wide barrier, normalisation: one launch) == bc_bn_stats followed by bc_ew_fused, bit for
    bit, and leaves its barrier counters at zero (repeated launches on one workspace)."""
    from blockcopy import _C

    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = (1.5 * torch.randn(N, C, H, W, device="cuda", generator=g) + 0.7).half().contiguous(memory_format=torch.channels_last)
    w = (torch.rand(C, device="cuda", generator=g) + 0.5) if affine else None
    sh = (torch.randn(C, device="cuda", generator=g) * 0.3) if affine else None
    ws = torch.zeros(_C.BN_STATS_WORKSPACE, dtype=torch.uint8, device="cuda")
    mean, invstd = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    _C.bn_stats(x, mean, invstd, 1e-5, ws)
    want = torch.empty_like(x)
    _C.ew_fused(want, x, None, (mean, invstd, w, sh), relu=relu)
    for _ in range(3):
        m2, i2 = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        got = torch.full_like(x, float("nan"))
        _C.bn_norm(got, x, m2, i2, w, sh, 1e-5, relu, ws)
        assert torch.equal(m2, mean) and torch.equal(i2, invstd)
        assert torch.equal(got, want)
        assert int(ws[:8].view(torch.int32).abs().sum()) == 0
