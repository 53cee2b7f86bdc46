"""SwiftNet-RN18 as a CONSUMER of the blockcopy API -- the benchmark workload named by
BASELINE.json (configs 1, 3, 4, 5).

Architecture and, importantly, the SEQUENCE of torch calls are those of the reference's
semantic_segmentation/lib/models/swiftnet (swiftnet.py:13-97, util.py:40-138,
backbones/resnet.py:59-304): plain nn.Modules that know nothing about blocks except the single
``@blockcopy_noblocks`` on the pyramid-pooling module.  Parameter names are identical, so a
reference state_dict loads with strict=True (used by the parity tests to run both
implementations on the same weights).  The reference's own files run unchanged on this package
too (tests/test_reference_models.py, when baseline/_ref is staged).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from blockcopy import blockcopy_noblocks


def _bilinear(x, size):
    return F.interpolate(x, size, mode="bilinear")


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, cout, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        y += x if self.downsample is None else self.downsample(x)
        return self.relu(y)


class ResNet18Encoder(nn.Module):
    """torchvision-style ResNet-18 trunk returning the four pyramid levels (strides 4..32)."""

    def __init__(self):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.block_features = []
        for i, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            setattr(self, f"layer{i}", self._stage(planes, 2, stride))
            self.block_features.append(planes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _stage(self, planes, blocks, stride):
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        seq = [BasicBlock(self.inplanes, planes, stride, down)]
        self.inplanes = planes
        seq += [BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def forward_down(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        feats = []
        for i in range(1, 5):
            x = getattr(self, f"layer{i}")(x)
            feats.append(x)
        return feats


class _BNReluConv(nn.Sequential):
    """Pre-activation unit: BatchNorm -> ReLU -> conv kxk (padding k//2)."""

    def __init__(self, cin, cout, k=3, batch_norm=True, bias=False):
        super().__init__()
        if batch_norm:
            self.add_module("norm", nn.BatchNorm2d(cin))
        self.add_module("relu", nn.ReLU(inplace=False))
        self.add_module("conv", nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, bias=bias))


class _Upsample(nn.Module):
    """Decoder step: 1x1 bottleneck on the skip, bilinear x2 of the coarse map, add, blend 3x3."""

    def __init__(self, cin, skip_in, cout, k=3):
        super().__init__()
        self.bottleneck = _BNReluConv(skip_in, cin, k=1)
        self.blend_conv = _BNReluConv(cin, cout, k=k)

    def forward(self, x, skip):
        skip = self.bottleneck(skip)
        x = _bilinear(x, (x.shape[2] * 2, x.shape[3] * 2))
        x += skip
        return self.blend_conv(x)


class SpatialPyramidPooling(nn.Module):
    """Runs DENSELY over the combined plane (global average pools cannot run per block)."""

    def __init__(self, cin, num_levels=3, bt_size=128, level_size=42, out_size=128, grids=(8, 4, 2, 1)):
        super().__init__()
        self.grids = grids
        self.spp = nn.Sequential()
        self.spp.add_module("spp_bn", _BNReluConv(cin, bt_size, k=1))
        width = bt_size
        for i in range(num_levels):
            width += level_size
            self.spp.add_module(f"spp{i}", _BNReluConv(bt_size, level_size, k=1))
        self.spp.add_module("spp_fuse", _BNReluConv(width, out_size, k=1))

    @blockcopy_noblocks
    def forward(self, x):
        size = x.size()[2:4]
        aspect = size[1] / size[0]
        x = self.spp[0](x)
        levels = [x]
        for i in range(1, len(self.spp) - 1):
            g = self.grids[i - 1]
            pooled = F.adaptive_avg_pool2d(x, (g, max(1, round(aspect * g))))
            levels.append(_bilinear(self.spp[i](pooled), size))
        return self.spp[-1](torch.cat(levels, 1))


class SwiftNetRN18(nn.Module):
    def __init__(self, num_classes=19, num_features=128):
        super().__init__()
        self.backbone = ResNet18Encoder()
        self.num_classes = num_classes
        f = self.backbone.block_features
        self.spp = SpatialPyramidPooling(f[3], bt_size=num_features, level_size=num_features // 3,
                                         out_size=num_features)
        ups = [_Upsample(num_features, f[i], num_features) for i in range(3)]
        self.upsample = nn.ModuleList(list(reversed(ups)))
        self.logits = _BNReluConv(num_features, num_classes, k=1, bias=True)

    def forward(self, image):
        feats = self.backbone.forward_down(image)[::-1]
        x = self.spp(feats[0])
        for skip, up in zip(feats[1:], self.upsample):
            x = up(x, skip)
        return self.logits(x)


def fuse_conv_bn_(model: nn.Module) -> nn.Module:
    """Fold every eval-mode BatchNorm2d that directly follows a Conv2d inside the same parent into
    that conv (weight scale + bias), leaving nn.Identity behind -- what the reference driver does
    with lib/utils/bn_fusion.py before timing (test_swiftnet.py:113-115).  Pre-activation
    BN->ReLU->conv units are untouched, as there."""
    for parent in model.modules():
        prev = None
        for name, child in list(parent.named_children()):
            if isinstance(child, nn.BatchNorm2d) and isinstance(prev, nn.Conv2d) and not child.training:
                with torch.no_grad():
                    scale = child.weight / torch.sqrt(child.running_var + child.eps)
                    shift = child.bias - child.running_mean * scale
                    prev.weight.mul_(scale.reshape(-1, 1, 1, 1))
                    if prev.bias is None:
                        prev.bias = nn.Parameter(shift.clone())
                    else:
                        prev.bias.mul_(scale).add_(shift)
                setattr(parent, name, nn.Identity())
                child = None
            prev = child
    return model


def build_swiftnet_rn18(seed: int = 0, num_classes: int = 19, fuse_bn: bool = True) -> SwiftNetRN18:
    """Random-init (seeded), eval-mode, optionally BN-fused SwiftNet-RN18."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = SwiftNetRN18(num_classes=num_classes).eval()
    torch.random.set_rng_state(gen_state)
    return fuse_conv_bn_(model) if fuse_bn else model
