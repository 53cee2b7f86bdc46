"""Per-parameter gradient error of policy/fused_train.py against fp32 torch autograd (max-abs / max and relative L2)."""
import copy
import sys

import torch

sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
sys.path.insert(0, "tests")
from test_gpu_policy_train import _policy_net  # noqa: E402

from blockcopy.policy.fused_train import FusedPolicyTrainer  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
N, H, W = 1, 256, 512
net = _policy_net(3)
ref = copy.deepcopy(net)
ref16 = copy.deepcopy(net)
tr = FusedPolicyTrainer(net)
g = torch.Generator(device="cuda").manual_seed(N + H)
x = torch.randn(N, 26, H, W, device="cuda", generator=g)
R = torch.randn(N, 1, H // 32, W // 32, device="cuda", generator=g)
logits = tr.run_train(lambda x16: x16[:, :26].copy_(x), (N, 26, H, W), x.device)
(logits * R).mean().backward()
want = ref.layers(ref.backbone(x))
(want * R).mean().backward()
# a second reference: torch autocast fp16 (cuDNN fp16 activations), to see what fp16 storage alone costs
with torch.autocast("cuda", dtype=torch.float16):
    w16 = ref16.layers(ref16.backbone(x))
(w16.float() * R).mean().backward()
print("logits err", float((logits - want).abs().max()), "range", float(want.abs().max()), "| autocast err", float((w16.float() - want).abs().max()))
for (name, q), (_, r), (_, a) in zip(net.named_parameters(), ref.named_parameters(), ref16.named_parameters()):
    if r.grad is None:
        continue
    d = q.grad - r.grad
    da = a.grad - r.grad
    print(f"{name:40s} max-rel {float(d.abs().max() / r.grad.abs().max()):.4f}  l2-rel {float(d.norm() / r.grad.norm()):.4f}"
          f"  cos {float(torch.nn.functional.cosine_similarity(q.grad.flatten(), r.grad.flatten(), dim=0)):.5f}"
          f"  | autocast max-rel {float(da.abs().max() / r.grad.abs().max()):.4f} l2-rel {float(da.norm() / r.grad.norm()):.4f}")
