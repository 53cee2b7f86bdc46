"""A stand-in for the SECOND consumer of the blockcopy API, Pedestron's CSPBlockCopy
(Pedestron/mmdet/models/detectors/csp_blockcopy.py:46-95, necks/csp_neck.py:37-83,
anchor_heads/csp_head.py:130-152), which cannot run here (needs mmcv + compiled ops).  It uses the
same API calls and the same op set on blocks: the detector inlines the per-frame state machine
(to_tensorwrapper / process_temporal_features / to_blocks / combine_), the backbone has a dilated
3x3 conv (padding 2), the neck runs ConvTranspose2d per block + channel L2 norm + cat, the head runs
conv + GroupNorm + ReLU on blocks, then ``blockcopy.to_tensor`` and a dense conv.
Imports only ``blockcopy`` -- it runs on the reference package and on this one alike."""
import blockcopy
import torch
import torch.nn as nn


class L2Norm(nn.Module):
    def __init__(self, c, scale):
        super().__init__()
        self.weight = nn.Parameter(torch.full((c,), float(scale)))
        self.eps = 1e-10

    def forward(self, x):
        norm = x.pow(2).sum(dim=1, keepdim=True).sqrt() + self.eps
        return self.weight.unsqueeze(0).unsqueeze(2).unsqueeze(3).expand_as(x) * x / norm


class StandinDetector(nn.Module):
    """wide=False: small channel counts (32/64/16, a 3x3 stem).  wide=True: the channel structure of the real CSP
    (ResNet 7x7 stem, every count a multiple of 64, 192-channel concatenation, GroupNorm(8, 64)), at which this repo
    runs every block op on its own kernels."""

    def __init__(self, settings, wide=False):
        super().__init__()
        self.is_blockcopy_manager = True
        c1, c2, cu, ch, g = (64, 128, 64, 64, 8) if wide else (32, 64, 16, 32, 4)
        self.stem = nn.Conv2d(3, c1, 7, 2, 3) if wide else nn.Conv2d(3, c1, 3, 2, 1)
        self.layer = nn.Conv2d(c1, 64, 3, 2, 1)
        self.dil = nn.Conv2d(64, 64, 3, 1, padding=2, dilation=2)
        self.down = nn.Conv2d(64, c2, 3, 2, 1)
        self.p_a = nn.ConvTranspose2d(64, cu, kernel_size=4, stride=2, padding=1)
        self.p_b = nn.ConvTranspose2d(c2, cu, kernel_size=4, stride=4, padding=0)
        self.l2_a, self.l2_b = L2Norm(cu, 10), L2Norm(cu, 10)
        self.head_conv = nn.Conv2d(c1 + 2 * cu, ch, 3, padding=1)
        self.head_gn = nn.GroupNorm(g, ch)
        self.cls = nn.Conv2d(ch, 8, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.policy = blockcopy.build_policy_from_settings(settings)
        self.train_interval = settings["block_train_interval"]
        self.block_temporal_features = None
        self.reset_temporal()

    def reset_temporal(self):
        self.clip_length = 0
        if self.block_temporal_features:
            self.block_temporal_features.clear()
        self.block_temporal_features = None
        self.policy_meta = {"inputs": None, "outputs": None, "outputs_prev": None}

    def extract_feat(self, x):
        f1 = self.relu(self.stem(x))
        f2 = self.relu(self.dil(self.relu(self.layer(f1))))
        f3 = self.relu(self.down(f2))
        return torch.cat([f1, self.l2_a(self.p_a(f2)), self.l2_b(self.p_b(f3))], dim=1)

    def bbox_head(self, x):
        feat = self.relu(self.head_gn(self.head_conv(x)))
        feat = blockcopy.to_tensor(feat)
        return self.cls(feat)

    def simple_test(self, img):
        self.clip_length += 1
        self.policy_meta["inputs"] = img
        self.policy_meta = self.policy(self.policy_meta)
        if self.policy_meta["num_exec"] == 0:
            self.policy_meta = self.policy_meta.copy()
            out = self.policy_meta["outputs"]
        else:
            x = blockcopy.to_tensorwrapper(img)
            self.block_temporal_features = x.process_temporal_features(self.block_temporal_features)
            x = x.to_blocks(self.policy_meta["grid"])
            self.policy_meta["frame_state"] = x.combine_().to_tensor()
            out = self.bbox_head(self.extract_feat(x))
        self.policy_meta["outputs_prev"] = self.policy_meta["outputs"]
        self.policy_meta["outputs"] = out
        self.policy_meta = self.policy.optim(self.policy_meta, train=self.clip_length % self.train_interval == 0)
        return out
