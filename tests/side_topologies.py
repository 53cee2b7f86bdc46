"""A small CNN built to hit the code paths of the deferred-launch engine that SwiftNet does not (used by the GPU
graph-vs-eager test and by the CPU lazy-vs-op-by-op test)."""
import torch


class SideTopologies(torch.nn.Module):
    """Shapes the side-stream logic must survive besides SwiftNet's: 1x1 convs IN the main chain (bottleneck),
    a pre-activation 1x1 unit whose consumer is a padded conv (its producer must then write the plane on the main
    stream), a side result with two consumers, and an in-place op on a tensor a side kernel reads."""

    def __init__(self):
        import torch.nn as nn

        super().__init__()
        self.stem = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.a1, self.a2, self.a3 = nn.Conv2d(64, 64, 1, bias=False), nn.Conv2d(64, 64, 3, 1, 1, bias=False), nn.Conv2d(64, 128, 1, bias=False)
        self.ds = nn.Conv2d(64, 128, 1, bias=False)
        self.bn_s, self.skip = nn.BatchNorm2d(128), nn.Conv2d(128, 64, 1, bias=False)
        self.bn_p, self.pre = nn.BatchNorm2d(128), nn.Conv2d(128, 64, 1, bias=False)
        self.after_pre = nn.Conv2d(64, 64, 3, 1, 1, bias=False)
        self.out = nn.Conv2d(64, 64, 3, 1, 1, bias=False)
        self.down = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
        self.blend = nn.Conv2d(64, 64, 3, 1, 1, bias=False)
        self.blend2 = nn.Conv2d(64, 64, 3, 1, 1, bias=False)
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                torch.nn.init.uniform_(m.weight, 0.5, 1.5)
                torch.nn.init.uniform_(m.bias, -0.2, 0.2)
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)

    def forward(self, x):
        import torch.nn.functional as F

        x = F.relu(self.stem(x))
        y = F.relu(self.a1(x))              # 1x1 on the main chain, consumed by a padded conv
        y = F.relu(self.a2(y))
        y = self.a3(y)                      # 1x1 on the main chain, consumed by the residual add
        y += self.ds(x)                     # residual downsample: side branch
        y = F.relu(y)
        s = self.skip(F.relu(self.bn_s(y)))  # pre-activation 1x1 unit: side branch, two consumers below
        p = self.pre(F.relu(self.bn_p(y)))   # pre-activation 1x1 unit consumed by a padded conv
        p = F.relu(self.after_pre(p))
        z = p + s
        z = z * 0.5                         # generic torch op on blocks (materialises, reads side results)
        s.mul_(2.0)                         # in-place op on a side result
        o = self.out(z + s)
        # upsample + add written into the consumer's plane only; a second padded consumer and a plain reader follow
        u = F.interpolate(self.down(o), scale_factor=2, mode="bilinear", align_corners=False)
        u += o
        return self.blend(u) + self.blend2(u) + u
