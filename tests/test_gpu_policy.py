"""bc_policy_features / bc_info_gain against the torch op sequences of the reference
(policy/net.py:84-113, policy/information_gain.py:32-41)."""
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
